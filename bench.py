"""Benchmark of the RVC synthesizer decode hot path (Synthesizer.infer).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config v2-48k]

Metric (BASELINE.json): audio-seconds synthesized per second at the model's native rate.

Workload (BASELINE.json configs[1]): RVC v2 48 kHz decode of one 60 s clip, cut by the
pipeline's silence split (x_pad, x_query, x_center, x_max = 1, 6, 38, 41 s --
rvc/infer/infer.py:41-45, rvc/infer/pipeline.py:330-348) into overlapping segments; one "step"
decodes every segment of the clip (B = 1 per segment, as pipeline.py:381-447 does).  Synthetic
features (phone ~ N(0,1), log-swept F0 with an unvoiced gap) and random-init weights of the named
architecture.  With --gpus N every rank decodes its own clip per step (segments are independent:
weak scaling, no collective in the decode); value = audio-seconds of all ranks / max-over-ranks
device time.

--impl reference times the reference algorithm's CPU path (the oracle port of
Synthesizer.infer: the same torch fp32 ATen conv kernels the reference dispatches to) on the
box's host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio-seconds synthesized/sec @48 kHz (1/2/4/8 B200) vs host-CPU reference"
UNIT = "audio-s/s"
CLIP_SECONDS = 60


def clip_segments(cfg, seed=0):
    """60 s synthetic clip -> the frame counts of its silence-split segments (incl. x_pad overlap)."""
    import numpy as np
    from polgen_rvc_b200 import segments as seg
    rng = np.random.default_rng(seed)
    n = seg.SAMPLE_RATE * CLIP_SECONDS
    t = np.arange(n) / seg.SAMPLE_RATE
    env = 0.15 + np.abs(np.sin(2 * np.pi * t / 7.3)) * (0.5 + 0.5 * np.sin(2 * np.pi * t / 1.7) ** 2)
    audio = rng.standard_normal(n) * env
    plan = seg.SegmentPlan()
    cuts = seg.split_points(audio, plan)
    return [c for _, c in seg.segment_frames(n, cuts, plan)]


# DRAM traffic of the top-time decoder conv shape from one `ncu --set full` capture (bench.py cannot run
# ncu on itself): C=128 k=11 conv1 at T=3435 (L = 412 200 rows), dram__bytes_read + dram__bytes_write.
NCU_TRAFFIC = {"kernel": "conv_planes_kernel<2,4,1> C=128 k=11 dil=1, L=412200", "bytes_per_launch": 160.32e6,
               "algorithmic_bytes_per_launch": 2 * 412200 * (128 + 128),
               "source": "profiles/r01p_ncu_full_conv_planes_c128k11.txt (launch 0; kernel unchanged since)"}

CPU_SAMPLE_FRAMES = 1000     # BASELINE.json's 10 s row: the CPU arms time one such segment per step


def workload_name(config, seg_frames):
    """the one workload both arms are quoted on (BASELINE.json configs[1])"""
    return (f"RVC {config} synthesizer decode, {CLIP_SECONDS} s clip split into silence segments of "
            f"{seg_frames} frames (incl. 1 s pad each side), B=1 per segment, one clip per GPU per step")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        if os.environ.get("PG_BENCH_SAMPLER_MS", "50") == "0":
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", os.environ.get("PG_BENCH_SAMPLER_MS", "50"),
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 - 0.05 <= ts <= t1 + 0.15 and len(r) >= 8] or \
               [r for _, r in self.rows if len(r) >= 8]
        if not rows:
            return None
        def num(v):
            try:
                return float(v)
            except ValueError:
                return float("nan")
        sm = [num(r[1]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": num(rows[0][2]),
                "power_w_max": max(num(r[3]) for r in rows), "samples": len(rows), "reasons": reasons}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1405.0), d.get("hbm_gbs", 6457.4), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_rate(cfg, frames, steps, warmup, seed=0):
    """Oracle port of Synthesizer.infer on the host cores: audio-s/s on one T=frames segment."""
    import torch
    import polgen_rvc_b200 as pg
    from oracle import rvc_oracle as orc
    # all the host cores this process may use (torchrun pins OMP_NUM_THREADS=1 for N>1 ranks)
    try:
        n_threads = len(os.sched_getaffinity(0))
    except AttributeError:
        n_threads = os.cpu_count() or 1
    torch.set_num_threads(max(1, n_threads))
    sd = orc.fold_weight_norm(pg.synth_weights(cfg, seed=seed))
    phone, lengths, pitch, f0, sid = pg.synth_inputs(cfg, 1, frames, seed=seed)
    eps_zp, eps_src = pg.synth_noise(cfg, 1, frames, seed=seed)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc.infer(sd, cfg, phone, lengths, pitch, f0, sid, eps_zp, eps_src, folded=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return frames / 100.0 / (sum(times) / len(times)), sum(times) / len(times), torch.get_num_threads()


def run_reference(args):
    """--impl reference: rank 0 alone times the CPU path; other ranks exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    import polgen_rvc_b200 as pg
    cfg = pg.CONFIGS[args.config]
    frames = CPU_SAMPLE_FRAMES     # bounded sample: one 10 s segment of the same config per step
    rate, sec, threads = cpu_reference_rate(cfg, frames, args.steps, args.warmup)
    sample = f"one {frames / 100:.0f} s segment (T={frames}, B=1) of {args.config} per step, torch fp32 CPU"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, clip_segments(cfg, seed=0)), "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host": {"cpu_count": os.cpu_count(), "torch_threads": threads, "torch": torch.__version__},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import polgen_rvc_b200 as pg
    from polgen_rvc_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # the sampler process lives for the whole run: nvidia-smi attaching to / detaching from the driver
    # stalls CUDA calls of this process for tens to hundreds of ms, so neither may fall near a timed region
    sampler = ClockSampler(local)
    sampler.start()

    cfg = pg.CONFIGS[args.config]
    seg_frames = clip_segments(cfg, seed=rank)
    sd = pg.synth_weights(cfg, seed=0)
    folded = pg.fold_state_dict(sd)
    # the product path: segments of the clip dealt over `lanes` engines / streams on this GPU
    sched = pg.SegmentScheduler(cfg, folded, local, lanes=args.lanes)
    # a profiled single engine for the per-kernel roofline pass (sequential, so CUDA-event durations
    # of one kernel are not inflated by another lane's kernels sharing the SMs)
    eng = pg.Engine(cfg, folded, local, _lib.PG_FLAG_PROFILE)
    segs_dev, segs_host, waves_host, waves_dev = [], [], [], []
    for i, T in enumerate(seg_frames):
        inp = pg.synth_inputs(cfg, 1, T, seed=100 * rank + i)
        segs_host.append([t.pin_memory() for t in inp])
        segs_dev.append([t.to(dev) for t in inp])
        waves_host.append(torch.empty(1, T * cfg.upp, dtype=torch.float32).pin_memory())
        waves_dev.append(torch.empty(1, T * cfg.upp, dtype=torch.float32, device=dev))
    audio_s = sum(seg_frames) / 100.0
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_device():
        return sched.decode(segs_dev, out=waves_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)        # also warms the fill kernel (lazy module load would land in step 1)
        step_device()
    torch.cuda.synchronize()

    # ---- timed region: K steps back to back (a stream of clips: the scheduler's lanes run on from one
    # step into the next, so the encoder phase of one clip overlaps the decoder phase of another);
    # device time by CUDA events around the K steps, L2 flushed once per step
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_cpu0 = time.perf_counter()
    marks, host_marks = [], []
    for _ in range(args.steps):
        flush.fill_(1)
        sched.decode(segs_dev, out=waves_dev, join=not args.pipelined)
        if os.environ.get("PG_BENCH_TRACE"):
            sched.join()
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
            host_marks.append(round((time.perf_counter() - t_cpu0) * 1e3, 2))
    sched.join()
    e1.record()
    host_enqueue_ms = (time.perf_counter() - t_cpu0) * 1e3 / args.steps
    if marks:
        torch.cuda.synchronize()
        print("device ms at end of each step:", [round(e0.elapsed_time(m), 2) for m in marks],
              "\nhost ms when each step was enqueued:", host_marks, file=sys.stderr)
    barrier()
    t_wall1 = time.time()
    dev_ms = e0.elapsed_time(e1)
    per_step_launches = sched.launch_count()

    # ---- roofline pass: the same segments, sequentially on the profiled engine (L2 flushed per step)
    for s in segs_dev:
        eng.infer(*s, None, None, 0, want_aux=False)
    torch.cuda.synchronize()
    eng.profile_read()
    eng.profile_table()     # drop the warm-up records
    seq_evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in segs_dev:
            eng.infer(*s, None, None, 0, want_aux=False)
        e1.record()
        seq_evs.append((e0, e1))
    torch.cuda.synchronize()
    seq_ms = sum(a.elapsed_time(b) for a, b in seq_evs)
    prof = eng.profile_read()
    table = eng.profile_table()

    # ---- e2e: the public scheduler call with HOST (pinned) buffers, H2D + D2H inside the timed region
    sched.decode(segs_host, host_out=waves_host)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    trace = []
    for _ in range(args.steps):
        t_s = time.perf_counter()
        sched.decode(segs_host, host_out=waves_host, join=not args.pipelined)
        trace.append(round((time.perf_counter() - t_s) * 1e3, 2))
    sched.join(host_sync=True)
    e1.record()
    if os.environ.get("PG_BENCH_TRACE"):
        print("e2e host ms per step:", trace, file=sys.stderr)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    h2d = sum(sum(t.numel() * t.element_size() for t in s) for s in segs_host)
    d2h = sum(w.numel() * 4 for w in waves_host)

    clocks = sampler.stop(t_wall0, t_wall1)
    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = float(t[0]), float(t[1])
    total_audio = audio_s * world * args.steps

    if rank == 0:
        tf_peak, hbm_peak, peak_src = peaks()
        umma_ms, umma_fl, umma_n = prof["conv_planes"]
        achieved = umma_fl / (umma_ms * 1e-3) / 1e12 if umma_ms > 0 else 0.0
        roofline = {
            "bound": "tensor", "kernel": "decoder tcgen05/TMEM implicit-GEMM convs over channel planes: conv_planes_kernel (ResBlock convs, ups, conv_pre) + pair_planes_kernel (fused ResBlock pairs, C=32/64)",
            "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
            "peak_source": peak_src, "traffic": NCU_TRAFFIC["bytes_per_launch"], "traffic_detail": NCU_TRAFFIC,
            "launches": umma_n, "avg_launch_ms": umma_ms / max(umma_n, 1),
            "algorithmic_flops_per_launch": umma_fl / max(umma_n, 1),
            "share_of_step": umma_ms / seq_ms if seq_ms > 0 else None,
            "measured_in": "sequential single-engine pass after the timed region (CUDA events around every "
                           "launch on its stream); the timed region itself overlaps segments on %d lanes" % args.lanes,
            "sequential_ms_per_step": seq_ms / args.steps,
        }
        line = {
            "metric": METRIC, "value": total_audio / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": workload_name(args.config, seg_frames),
                       "audio_s_per_step_per_gpu": audio_s, "l2": "flushed between steps (256 MiB write)",
                       "segment_lanes_per_gpu": args.lanes,
                       "steps_pipelined": bool(args.pipelined),
                       "parallelism": f"segment-sharded x{world}, no collective"},
            "roofline": roofline,
            "e2e": {"value": total_audio / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": per_step_launches * args.steps,
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "clocks": clocks,
            "generator_tflops": cfg.generator_flops_per_frame() * 100 * total_audio / (dev_ms_max * 1e-3) / 1e12,
        }
        if args.table:
            # per-layer-shape breakdown of the tensor-core conv launches inside the timed region
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", args.table), "w") as f:
                f.write("class,Cin,N,K,dil,launches,ms_total,ms_per_launch,tflops,share_of_step\n")
                for cls, cin, n, k, dil, cnt, ms, fl in sorted(table, key=lambda r: -r[6]):
                    f.write(f"{('planes-tcgen05', 'cuda-core', 'nlc-tcgen05-split')[int(cls)]},{int(cin)},{int(n)},{int(k)},{int(dil)},{int(cnt)},"
                            f"{ms:.3f},{ms / cnt:.4f},{fl / (ms * 1e-3) / 1e12 if ms > 0 else 0:.1f},{ms / seq_ms:.4f}\n")
        if world == 1 and not args.no_cpu:
            frames = CPU_SAMPLE_FRAMES
            rate, sec, threads = cpu_reference_rate(cfg, frames, 6, 1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"one {frames / 100:.0f} s segment (T={frames}, B=1) of {args.config}, "
                                              f"oracle port of Synthesizer.infer, torch fp32, mean of 6 after 1 warm-up"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="v2-48k", choices=["v2-48k", "v2-40k", "v2-32k", "v1-40k"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pipelined", dest="pipelined", action="store_false",
                    help="join the scheduler's lanes after every step instead of streaming clip after clip")
    ap.add_argument("--lanes", type=int, default=2, help="engines/streams per GPU the clip's segments are dealt over")
    ap.add_argument("--table", default="", help="write the per-layer-shape conv timing table to gpurun_out/<name>")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
