"""TextEncoder attention A/B on the bench clip's ragged batch: tcgen05 / TMEM kernel (default) against the
mma.sync twin (PG_FLAG_LEGACY_ATTENTION).  Prints the TextEncoder time of both (CUDA events; everything but the
attention is identical) and the latent agreement.

    python tools/attn_check.py [--T 3435,2965] [--iters 20]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", default="3435,2965")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--config", default="v2-48k")
    args = ap.parse_args()
    import torch
    import polgen_rvc_b200 as pg
    from polgen_rvc_b200 import _lib
    lens = [int(x) for x in args.T.split(",")]
    B, T = len(lens), max(lens)
    cfg = pg.CONFIGS[args.config]
    folded = pg.fold_state_dict(pg.synth_weights(cfg, seed=0))
    d = torch.device("cuda", 0)
    phone, lengths, pitch, f0, sid = pg.synth_inputs(cfg, B, T, seed=0)
    lengths = torch.tensor(lens)
    phone, lengths, pitch = phone.to(d), lengths.to(d), pitch.to(d)
    res = {}
    for name, flags in (("tcgen05", 0), ("mma_sync_twin", _lib.PG_FLAG_LEGACY_ATTENTION)):
        eng = pg.Engine(cfg, folded, 0, flags)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(3):
                m, l = eng.text_encoder(phone, lengths, pitch)
            st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(args.iters):
                m, l = eng.text_encoder(phone, lengths, pitch)
            e1.record(st)
            st.synchronize()
        res[name] = {"text_encoder_ms": e0.elapsed_time(e1) / args.iters, "m": m.clone(), "l": l.clone()}
        eng.close()
    a, b = res["tcgen05"], res["mma_sync_twin"]
    scale = max(1.0, float(b["m"].abs().max()))
    print(json.dumps({
        "lens": lens,
        "text_encoder_ms_tcgen05": a["text_encoder_ms"], "text_encoder_ms_mma_sync": b["text_encoder_ms"],
        "attention_ms_saved_per_call": b["text_encoder_ms"] - a["text_encoder_ms"],
        "m_p_max_abs_diff_rel": float((a["m"] - b["m"]).abs().max()) / scale,
        "logs_p_max_abs_diff": float((a["l"] - b["l"]).abs().max()),
        "finite": bool(torch.isfinite(a["m"]).all() and torch.isfinite(a["l"]).all())}), flush=True)


if __name__ == "__main__":
    main()
