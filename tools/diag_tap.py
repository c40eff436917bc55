"""diagnostic: repeat a decode with taps kept; where does the last-stage mean differ between calls?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polgen_rvc_b200 as pg
from polgen_rvc_b200 import _lib
cfg = pg.CONFIGS["v2-40k"]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
folded = pg.fold_state_dict(pg.synth_weights(cfg, seed=0))
d = torch.device("cuda:0")
B, T = 2, 410
inp = [t.to(d) for t in pg.synth_inputs(cfg, B, T, seed=80)]
eng = pg.Engine(cfg, folded, 0, _lib.PG_FLAG_KEEP_TAPS)
ref = {}
names = ["dec.stage2", "dec.stage3"]
nbad = 0
for rep in range(reps):
    w, _ = eng.infer(*inp, None, None, 11, want_aux=False)
    torch.cuda.synchronize()
    taps = {n: eng.fetch_tap(n).clone() for n in names}
    if not ref:
        ref = taps
        continue
    for n in names:
        dif = (taps[n] != ref[n])
        if dif.any():
            nbad += 1
            idx = dif.nonzero()
            b0, t0, c0 = idx[0].tolist()
            rows = sorted(set(idx[:, 1].tolist()))
            chans = sorted(set(idx[:, 2].tolist()))
            dv = (taps[n] - ref[n]).abs()
            rel = float(dv.max() / ref[n].abs().max())
            print(f"rep {rep} {n}: {int(dif.sum())} elems, b={b0} rows {rows[0]}..{rows[-1]} ({len(rows)}) chans {chans[0]}..{chans[-1]} ({len(chans)}) "
                  f"maxdiff {float(dv.max()):.3e} (tensor max {float(ref[n].abs().max()):.3f}) at value {float(ref[n][b0, t0, c0]):.4f} -> {float(taps[n][b0, t0, c0]):.4f}")
            break
print("calls with differences:", nbad, "of", reps - 1)
