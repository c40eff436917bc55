"""Lane / decoder-SM-cap sweep on the bench clip (one process, one GPU):

    python tools/sweep_overlap.py [--steps 12] [--config v2-48k] [--grid "2:0,2:140,2:132,..."]

For every (lanes, decoder_sms) pair: a SegmentScheduler streams the bench's clip variants (bench.py's
device-resident `value` leg) and the ms per clip is printed.  Also checks that a capped decoder returns
bit-identical waveforms (the persistent kernels walk their tiles grid-stride; a tile's result does not
depend on the CTA that computes it).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--config", default="v2-48k")
    ap.add_argument("--grid", default="1:0,2:0,2:144,2:140,2:136,2:132,2:124,2:116,3:132,3:124")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import polgen_rvc_b200 as pg
    import bench

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cfg = pg.CONFIGS[args.config]
    folded = pg.fold_state_dict(pg.synth_weights(cfg, seed=0))
    variants = []
    for v in range(bench.N_VARIANTS):
        frames = bench.clip_variant(bench.clip_segments(cfg, seed=0), v)
        segs = [[t.to(dev) if j != 4 else t for j, t in enumerate(pg.synth_inputs(cfg, 1, T, seed=7 * v + i))]
                for i, T in enumerate(frames)]
        outs = [torch.empty(1, T * cfg.upp, dtype=torch.float32, device=dev) for T in frames]
        variants.append((frames, segs, outs))
    audio_s = sum(variants[0][0]) / 100.0
    ref_wave, rows = None, []
    for item in args.grid.split(","):
        lanes, sms = (int(x) for x in item.split(":"))
        sched = pg.SegmentScheduler(cfg, folded, 0, lanes=lanes, decoder_sms=sms)
        if lanes == 1 and sms:
            sched.engines[0].set_decoder_sms(sms)
        for i in range(max(args.warmup, 3)):
            frames, segs, outs = variants[i % len(variants)]
            sched.decode(segs, seed=i, out=outs, join=False)
        sched.join(host_sync=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            frames, segs, outs = variants[i % len(variants)]
            sched.decode(segs, seed=i, out=outs, join=False)
        sched.join()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        # determinism across caps: one clip, fixed seed
        frames, segs, outs = variants[0]
        sched.decode(segs, seed=1234, out=outs, join=True)
        torch.cuda.synchronize()
        wave = torch.cat([o.reshape(-1) for o in outs]).clone()
        same = None
        if ref_wave is None:
            ref_wave = wave
        else:
            same = bool(torch.equal(wave, ref_wave))
        row = {"lanes": lanes, "decoder_sms": sms, "ms_per_clip": ms, "audio_s_per_s": audio_s / (ms * 1e-3),
               "bit_identical_to_first": same}
        rows.append(row)
        print(json.dumps(row), flush=True)
        sched.close()
        del sched
        torch.cuda.empty_cache()
    if args.out:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", args.out), "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
