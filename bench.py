"""Benchmark of the RVC synthesizer decode hot path (Synthesizer.infer).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config v2-48k|v2-40k|v2-32k|v1-40k] [--workload clip|batch512]

Metric (BASELINE.json): audio-seconds synthesized per second at the model's native rate.

Workload "clip" (default; BASELINE.json configs[1]): decode of one 60 s clip per GPU per step, cut by
the pipeline's silence split (x_pad, x_query, x_center, x_max = 1, 6, 38, 41 s -- rvc/infer/infer.py:41-45,
rvc/infer/pipeline.py:330-348) into overlapping segments.  The product path decodes a clip's segments as
one ragged batch (pg_infer_segments) on one of two lanes per GPU.  Successive steps use clips whose
segment lengths DIFFER (a real stream of clips never repeats a length), so CUDA-graph reuse comes from
the length buckets, not from replaying one shape.  With --gpus N the N clips' segments are bin-packed
over the ranks by polgen_rvc_b200.segments.plan_shards (weak scaling, no collective in the decode);
value = audio-seconds of all ranks / max-over-ranks device time.

Workload "batch512" (BASELINE.json configs[2]): 512 x 10 s segments, plan_shards over the ranks, decoded
in equal-length sub-batches of 64, waveforms gathered on rank 0 (strong scaling: total work fixed).

Synthetic features (phone ~ N(0,1), log-swept F0 with an unvoiced gap) and random-init weights of the
named architecture.

--impl reference times the reference's CPU path on the box's host cores on a bounded sample of the same
config: the UNMODIFIED reference module from baseline/_ref when __graft_entry__.build() could vendor it
(kind "reference"), else the oracle port (kind "port").
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
CLIP_SECONDS = int(os.environ.get("PG_BENCH_CLIP_SECONDS", "60"))   # BASELINE configs[1]: 60 s (the override is a tuning aid)
CPU_SAMPLE_FRAMES = 1000     # BASELINE.json's 10 s row: the CPU arms time one such segment per step
N_VARIANTS = 4               # clips with different segment lengths cycled through the timed steps
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


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


def clip_variant(frames, v):
    """the cut point of variant v moved by a few frames: same clip length, different segment lengths"""
    if len(frames) < 2 or v == 0:
        return list(frames)
    d = (37 * v) % 211 - 90
    out = list(frames)
    out[0] += d
    out[-1] -= d
    return out


def workload_name(config, seg_frames):
    """the one workload both arms are quoted on (BASELINE.json configs[1])"""
    return (f"RVC {config} synthesizer decode, {CLIP_SECONDS} s clip split into silence segments of "
            f"{seg_frames} frames (incl. 1 s pad each side), one clip per GPU per step")


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


def ncu_traffic():
    """DRAM bytes per launch of the top decoder conv, from this round's `ncu --set full` capture
    (written by tools/ncu_summary.py --traffic; bench.py cannot run ncu on itself)."""
    for name in ("r03_traffic.json", "r02_traffic.json"):     # newest capture first
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                return json.load(open(p))
            except ValueError:
                pass
    return None


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))   # torchrun pins OMP_NUM_THREADS=1 for N>1 ranks
    except AttributeError:
        return os.cpu_count() or 1


def reference_module_available():
    return os.path.isfile(os.path.join(REF_DIR, "rvc", "lib", "algorithm", "synthesizers.py"))


def build_reference_module(cfg, sd, device="cpu"):
    """the UNMODIFIED reference Synthesizer (vendored under baseline/_ref by __graft_entry__.build()),
    constructed as rvc/infer/infer.py:92-102 does"""
    import torch
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from rvc.lib.algorithm.synthesizers import Synthesizer
    net = Synthesizer(*cfg.ctor_args(), use_f0=1, input_dim=cfg.input_dim, is_half=False)
    del net.enc_q
    net.load_state_dict(sd, strict=False)
    return net.eval().to(device).float()


def cpu_reference_rate(cfg, frames, steps, warmup, seed=0, keep=None):
    """The reference's CPU path on the host cores: audio-s/s on one T=frames segment.  Returns
    (rate, seconds per call, threads, kind); `keep` (dict) receives the inputs / noise / waveform of the
    last call for the in-run parity check."""
    import torch
    import polgen_rvc_b200 as pg
    from oracle import rvc_oracle as orc
    n_threads = host_threads()
    torch.set_num_threads(n_threads)
    sd = pg.synth_weights(cfg, seed=seed)
    inputs = pg.synth_inputs(cfg, 1, frames, seed=seed)
    eps_zp, eps_src = pg.synth_noise(cfg, 1, frames, seed=seed)
    use_ref = reference_module_available() and keep is None
    if use_ref:
        net = build_reference_module(cfg, sd)
        call = lambda: net.infer(*inputs)[0]
    else:
        W = orc.fold_weight_norm(sd)
        call = lambda: orc.infer(W, cfg, *inputs, eps_zp, eps_src, folded=True)[0]
    times, out = [], None
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = call()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    if keep is not None:
        keep.update(sd=sd, inputs=inputs, noise=(eps_zp, eps_src), wave=out)
    sec = sum(times) / len(times)
    return frames / 100.0 / sec, sec, torch.get_num_threads(), "reference" if use_ref else "port"


def run_reference(args):
    """--impl reference: rank 0 alone times the CPU path; other ranks exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    import polgen_rvc_b200 as pg
    cfg = pg.CONFIGS[args.config]
    frames = CPU_SAMPLE_FRAMES     # bounded sample: one 10 s segment of the same config per step
    rate, sec, threads, kind = cpu_reference_rate(cfg, frames, args.steps, args.warmup)
    what = "unmodified reference module (baseline/_ref)" if kind == "reference" else "oracle port"
    sample = f"one {frames / 100:.0f} s segment (T={frames}, B=1) of {args.config} per step, {what}, torch fp32 CPU"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, clip_segments(cfg, seed=0)), "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host": {"cpu_count": os.cpu_count(), "torch_threads": threads, "torch": torch.__version__},
    }
    print(json.dumps(line), flush=True)


def torch_eager_gpu(cfg, dev, frames=CPU_SAMPLE_FRAMES, iters=5):
    """the unmodified reference module run by torch eager (cuDNN/cuBLAS fp32) on this
    same GPU -- the existing Blackwell code path the hand-written kernels have to beat."""
    import torch
    import polgen_rvc_b200 as pg
    if not reference_module_available():
        return {"unavailable": "baseline/_ref not vendored (no /root/reference at build time)"}
    try:
        net = build_reference_module(cfg, pg.synth_weights(cfg, seed=0), dev)
        inputs = [t.to(dev) for t in pg.synth_inputs(cfg, 1, frames, seed=0)]
        res = {}
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                for _ in range(2):
                    net.infer(*inputs)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    net.infer(*inputs)[0][0, 0].data.cpu().float().numpy()
                e1.record()
                torch.cuda.synchronize()
            res["tf32" if tf32 else "fp32"] = frames / 100.0 * iters / (e0.elapsed_time(e1) * 1e-3)
        torch.backends.cudnn.allow_tf32 = True
        return {"audio_s_per_s_fp32": res["fp32"], "audio_s_per_s_tf32": res["tf32"], "unit": UNIT,
                "sample": f"{iters} x one {frames / 100:.0f} s segment (B=1), unmodified reference module, torch "
                          f"{torch.__version__} eager on this GPU, waveform to host each call"}
    except Exception as e:   # noqa: BLE001 -- evidence leg only: never fail the bench line
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import polgen_rvc_b200 as pg
    from polgen_rvc_b200 import _lib
    from polgen_rvc_b200 import segments as seg

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
    sd = pg.synth_weights(cfg, seed=0)
    folded = pg.fold_state_dict(sd)
    batch512 = args.workload == "batch512"
    sched = pg.SegmentScheduler(cfg, folded, local, lanes=args.lanes, max_batch=args.max_batch,
                                max_batch_frames=max(36000, args.max_batch * 1100),
                                decoder_sms=args.decoder_sms)
    conv = pg.ClipConverter(sched, cfg.sr, depth=max(2, args.lanes))
    warm = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the segment list of every step variant, sharded over the ranks by plan_shards
    variants = []      # per variant: dict(frames=[...local], host=[...], dev=[...], audio_s_global)
    if batch512:
        n_total, T = args.segments, CPU_SAMPLE_FRAMES
        bins = seg.plan_shards([T] * n_total, world)
        mine = bins[rank]
        base = [pg.synth_inputs(cfg, 1, T, seed=1000 + (i % 16)) for i in range(16)]   # 16 distinct segments, reused
        host = [[t.pin_memory() if torch.is_tensor(t) else t for t in base[i % 16]] for i in mine]
        variants.append({"frames": [T] * len(mine), "host": host, "bins": bins, "audio_s_global": n_total * T / 100.0})
        n_var = 1
    else:
        n_var = N_VARIANTS
        for v in range(n_var):
            all_frames, owner = [], []
            for r in range(world):
                fr = clip_variant(clip_segments(cfg, seed=r), v)
                all_frames += fr
                owner += [r] * len(fr)
            bins = seg.plan_shards(all_frames, world)
            mine = bins[rank]
            host = [[t.pin_memory() for t in pg.synth_inputs(cfg, 1, all_frames[i], seed=100 * owner[i] + 7 * v + i)]
                    for i in mine]
            variants.append({"frames": [all_frames[i] for i in mine], "host": host, "bins": bins,
                             "audio_s_global": sum(all_frames) / 100.0})
    for var in variants:
        # sid stays on the host (it travels as a kernel parameter); everything else is uploaded once
        var["dev"] = [[t.to(dev) if j != 4 else t for j, t in enumerate(s)] for s in var["host"]]
        var["out"] = [torch.empty(1, T * cfg.upp, dtype=torch.float32, device=dev) for T in var["frames"]]
    seg_frames0 = clip_segments(cfg, seed=0)

    # ---- value: inputs resident in HBM, K steps back to back (a stream of clips: the scheduler's lanes run
    # on from one step into the next).  No explicit L2 flush: one step streams > 3 GB of activations through
    # the 126 MB L2, so nothing of step i+1's inputs survives step i.
    for i in range(warm):
        var = variants[i % n_var]
        sched.decode(var["dev"], seed=i, out=var["out"], join=False)
    sched.join(host_sync=True)
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_cpu0 = time.perf_counter()
    launches = 0
    audio_value = 0.0
    for i in range(args.steps):
        var = variants[i % n_var]
        sched.decode(var["dev"], seed=i, out=var["out"], join=not args.pipelined)
        launches += sched.launch_count()
        audio_value += var["audio_s_global"]
    sched.join()
    e1.record()
    host_enqueue_ms = (time.perf_counter() - t_cpu0) * 1e3 / args.steps
    barrier()
    t_wall1 = time.time()
    dev_ms = e0.elapsed_time(e1)
    graphs = sum(e.graph_count() for e in sched.engines)

    # ---- e2e: the batched clip conversion (ClipConverter: pinned host features in -> ragged decode ->
    # trim/concat + peak normalise + int16 on the GPU -> pinned host int16 out), H2D and D2H inside the region
    def e2e_step(i):
        var = variants[i % n_var]
        if batch512:      # independent 10 s segments: decode into per-rank rows, gather on rank 0, D2H
            sched.decode(var["host"], seed=i, out=var["out"], join=True)
            rows = torch.cat(var["out"], 0)
            full = seg.gather_waveforms(rows, var["bins"], rows.shape[1], rank, world)
            if full is not None:
                host_full.copy_(full, non_blocking=True)
        else:
            conv.convert(var["host"], seed=i, sync=False)

    h2d = sum(sum(t.numel() * t.element_size() for t in s if torch.is_tensor(t)) for s in variants[0]["host"])
    if batch512:
        host_full = torch.empty(args.segments, CPU_SAMPLE_FRAMES * cfg.upp, dtype=torch.float32).pin_memory() if rank == 0 else None
        d2h = args.segments * CPU_SAMPLE_FRAMES * cfg.upp * 4 if rank == 0 else 0
    else:
        d2h = sum(T * cfg.upp - 2 * cfg.sr for T in variants[0]["frames"]) * 2
    for i in range(2):
        e2e_step(i)
    conv.wait()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_e2e0 = time.perf_counter()
    audio_e2e = 0.0
    for i in range(args.steps):
        e2e_step(i)
        audio_e2e += variants[i % n_var]["audio_s_global"]
    conv.wait()
    torch.cuda.synchronize()
    e2e_wall_ms = (time.perf_counter() - t_e2e0) * 1e3
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), e2e_wall_ms)     # host-visible completion: int16 is in pinned memory

    # ---- e2e through the UNCHANGED reference loop with the drop-in module (pipeline.py:272-286): one
    # net_g.infer per segment on torch's default stream, waveform pulled to numpy every call
    dropin = None
    if not batch512 and not args.no_dropin:
        net = pg.Synthesizer(*cfg.ctor_args(), use_f0=1, input_dim=cfg.input_dim, is_half=False)
        del net.enc_q
        net.load_state_dict(sd, strict=False)
        net.eval().to(dev)
        def dropin_step(i):
            n = 0
            for phone, lengths, pitch, f0, sid in variants[i % n_var]["host"]:
                with torch.no_grad():
                    audio1 = net.infer(phone.to(dev, non_blocking=True), lengths.to(dev), pitch.to(dev, non_blocking=True),
                                       f0.to(dev, non_blocking=True), sid.to(dev))[0][0, 0].data.cpu().float().numpy()
                n += audio1.shape[0]
            return n
        for i in range(2 * n_var):
            dropin_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            dropin_step(i)
        torch.cuda.synchronize()
        dropin_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        dropin = {"ms_per_step": dropin_ms / args.steps, "graphs": net.engine().graph_count()}
        del net

    # ---- roofline pass: the same ragged batches, sequentially on ONE profiled engine (CUDA events around
    # every conv launch on its stream; durations not inflated by another lane's kernels sharing the SMs)
    eng = pg.Engine(cfg, folded, local, _lib.PG_FLAG_PROFILE)
    def seq_pass(i):
        var = variants[i % n_var]
        for idx in sched.plan_batches(var["frames"]):
            eng.infer_segments([sched._segment_dict(var["dev"][k]) for k in idx], seed=i)
    seq_pass(0)
    torch.cuda.synchronize()
    eng.profile_read()
    eng.profile_table()     # drop the warm-up records
    n_seq = min(args.steps, 4) if batch512 else args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_seq):
        seq_pass(i)
    e1.record()
    torch.cuda.synchronize()
    seq_ms = e0.elapsed_time(e1)
    prof = eng.profile_read()
    table = eng.profile_table()
    eng.close()

    clocks = sampler.stop(t_wall0, t_wall1)
    t = torch.tensor([dev_ms, e2e_ms, dropin["ms_per_step"] if dropin else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, dropin_ms_max = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        tf_peak, hbm_peak, peak_src = peaks()
        umma_ms, umma_fl, umma_n = prof["conv_planes"]
        achieved = umma_fl / (umma_ms * 1e-3) / 1e12 if umma_ms > 0 else 0.0
        traffic = ncu_traffic()
        roofline = {
            "bound": "tensor", "kernel": "decoder tcgen05/TMEM implicit-GEMM convs over channel planes: conv_planes_kernel (ResBlock convs, ups, conv_pre) + pair_planes_kernel (fused ResBlock pairs, C=32/64)",
            "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
            "peak_source": peak_src, "traffic": traffic["bytes_per_launch"] if traffic else None,
            "traffic_detail": traffic,
            "launches": umma_n, "avg_launch_ms": umma_ms / max(umma_n, 1),
            "algorithmic_flops_per_launch": umma_fl / max(umma_n, 1),
            "share_of_step": umma_ms / seq_ms if seq_ms > 0 else None,
            "measured_in": "sequential single-engine pass after the timed region (CUDA events around every "
                           "conv launch on its stream); the timed region itself overlaps clips on %d lanes" % args.lanes,
            "sequential_ms_per_step": seq_ms / n_seq,
            # the five conv shapes with the largest share of the sequential step (class 0 = decoder plane convs; N < 0:
            # fused ResBlock pair), each with its own fraction of the peak: the class figure above is their FLOP-weighted mean
            "top_shapes": [
                {"Cin": int(cin), "N": int(n), "K": int(k), "dil": int(dil), "MT": int(mt), "launches": int(cnt),
                 "ms_per_launch": ms / cnt, "tflops": fl / (ms * 1e-3) / 1e12, "frac": fl / (ms * 1e-3) / 1e12 / tf_peak,
                 "share_of_step": ms / seq_ms}
                for cls, cin, n, k, dil, mt, cnt, ms, fl in sorted((r for r in table if int(r[0]) == 0 and r[7] > 0),
                                                                    key=lambda r: -r[7])[:5]],
        }
        wl = (f"RVC {args.config} synthesizer decode, {args.segments} x 10 s segments (T=1000) sharded over the ranks by "
              f"plan_shards, equal-length sub-batches of <= {args.max_batch}, waveforms gathered on rank 0"
              if batch512 else workload_name(args.config, seg_frames0))
        line = {
            "metric": METRIC, "value": audio_value / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong" if batch512 else "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": wl,
                       "audio_s_per_step": variants[0]["audio_s_global"],
                       "step_variants": [v["frames"] for v in variants] if not batch512 else None,
                       "l2": "no explicit flush: a step streams > 3 GB of activations through the 126 MB L2 (inputs larger than L2)",
                       "segment_lanes_per_gpu": args.lanes, "decoder_sms": args.decoder_sms or None, "ragged_batching": True,
                       "steps_pipelined": bool(args.pipelined),
                       "cuda_graphs_held": graphs,
                       "parallelism": f"segment-sharded x{world} (plan_shards), no collective in the decode"},
            "roofline": roofline,
            "e2e": {"value": audio_e2e / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "path": ("SegmentScheduler.decode (pinned host features) + NCCL gather of the waveforms to rank 0 + D2H (fp32)"
                             if batch512 else
                             "ClipConverter.convert: pinned host features -> pg_infer_segments (ragged batch) -> "
                             "pg_postprocess (trim/concat, peak normalise, int16) -> pinned host int16")},
            "gpu_launches": launches,
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "clocks": clocks,
            "generator_tflops": cfg.generator_flops_per_frame() * 100 * audio_value / (dev_ms_max * 1e-3) / 1e12,
        }
        if dropin:
            line["e2e_dropin"] = {
                "value": audio_e2e / (dropin_ms_max * args.steps * 1e-3), "unit": UNIT,
                "path": "the reference's own loop (pipeline.py:272-286) around the drop-in module: net_g.infer per "
                        "segment on torch's default stream, [0][0,0].data.cpu().float().numpy() per call, distinct "
                        "segment lengths every step", "cuda_graphs_held": dropin["graphs"]}
        if args.table:
            # per-layer-shape breakdown of the tensor-core conv launches of the roofline pass
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", args.table), "w") as f:
                f.write("class,Cin,N,K,dil,MT,launches,ms_total,ms_per_launch,tflops,share_of_step\n")
                for cls, cin, n, k, dil, mt, cnt, ms, fl in sorted(table, key=lambda r: -r[7]):
                    f.write(f"{('planes-tcgen05', 'cuda-core', 'nlc-tcgen05-split')[int(cls)]},{int(cin)},{int(n)},{int(k)},{int(dil)},{int(mt)},{int(cnt)},"
                            f"{ms:.3f},{ms / cnt:.4f},{fl / (ms * 1e-3) / 1e12 if ms > 0 else 0:.1f},{ms / seq_ms:.4f}\n")
        if world == 1 and not args.no_cpu:
            frames = CPU_SAMPLE_FRAMES
            keep = {}
            rate, sec, threads, kind = cpu_reference_rate(cfg, frames, 6, 1, keep=keep)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"one {frames / 100:.0f} s segment (T={frames}, B=1) of {args.config}, "
                                              f"oracle port of Synthesizer.infer, torch fp32, mean of 6 after 1 warm-up"}
            # in-run parity: the same segment, same weights and noise, through the product engine
            peng = pg.Engine(cfg, folded, local)
            ez, es = keep["noise"]
            w, _ = peng.infer(*[t.to(dev) for t in keep["inputs"]], ez.transpose(1, 2).contiguous().to(dev),
                              es.reshape(1, -1).contiguous().to(dev), 0, want_aux=False)
            torch.cuda.synchronize()
            want = keep["wave"][:, 0].double()
            err = w.cpu().double() - want
            line["parity"] = {"snr_db": float(10 * torch.log10((want ** 2).sum() / (err ** 2).sum())),
                              "max_abs": float(err.abs().max()), "T": frames,
                              "against": "oracle (fp32 CPU) on the cpu_baseline segment, same weights and noise",
                              "tolerance": "SNR >= 40 dB, max-abs <= 1e-2"}
            peng.close()
            if not args.no_eager:
                line["reference_gpu_eager"] = torch_eager_gpu(cfg, dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="v2-48k", choices=["v2-48k", "v2-40k", "v2-32k", "v1-40k"])
    ap.add_argument("--workload", default="clip", choices=["clip", "batch512"])
    ap.add_argument("--segments", type=int, default=512, help="batch512: number of 10 s segments")
    ap.add_argument("--max-batch", type=int, default=64, help="rows per ragged / equal-length sub-batch")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity / torch-eager legs")
    ap.add_argument("--no-eager", action="store_true", help="skip the torch-eager-on-GPU reference timing")
    ap.add_argument("--no-dropin", action="store_true", help="skip the e2e_dropin leg")
    ap.add_argument("--no-pipelined", dest="pipelined", action="store_false",
                    help="join the scheduler's lanes after every step instead of streaming clip after clip")
    ap.add_argument("--lanes", type=int, default=2, help="engines/streams per GPU the batches are dealt over")
    ap.add_argument("--decoder-sms", type=int, default=int(os.environ.get("PG_BENCH_DECODER_SMS", "0")),
                    help="cap on the CTAs of the decoder's persistent kernels (0 = one per SM); with >= 2 lanes the SMs left "
                         "free run the other lane's TextEncoder / flow kernels beside the decoder")
    ap.add_argument("--table", default="", help="write the per-layer-shape conv timing table to gpurun_out/<name>")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
