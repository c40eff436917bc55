"""Find the first decoder/encoder tap that differs between two identical calls separated by a call with
different noise (developer tool, run under gpurun)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import polgen_rvc_b200 as pg
from polgen_rvc_b200 import _lib

cfg = pg.CONFIGS["v2-48k"]
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
flags = _lib.PG_FLAG_KEEP_TAPS | (int(sys.argv[2]) if len(sys.argv) > 2 else 0)
sd = pg.synth_weights(cfg, seed=0)
eng = pg.Engine(cfg, pg.fold_state_dict(sd), 0, flags)
d = torch.device("cuda:0")
one = [t.to(d) for t in pg.synth_inputs(cfg, 1, T, seed=0)]
names = ["enc.x0"] + [f"enc.layer{i}" for i in range(6)] + [f"flow.{f}" for f in range(4)] + ["source", "dec.conv_pre"]
for i in range(4):
    names += [f"dec.ups{i}", f"dec.stage{i}"]

def run(seed):
    w, aux = eng.infer(*one, None, None, seed)
    torch.cuda.synchronize()
    taps = {n: eng.fetch_tap(n).clone() for n in names}
    taps["wave"] = w.clone()
    return taps

two = [torch.cat([t, t]) for t in one]
mode = sys.argv[4] if len(sys.argv) > 4 else "seed"
bad = 0
for trial in range(int(sys.argv[3]) if len(sys.argv) > 3 else 6):
    if mode == "b2":
        eng.infer(*two, None, None, 5 + trial)
        a = run(1)
        run(2 + trial)
        b = run(1)
    else:
        a = run(1)
        run(2 + trial)
        b = run(1)
    diffs = [(n, int((a[n] != b[n]).sum()), float((a[n] - b[n]).abs().max())) for n in list(names) + ["wave"] if not torch.equal(a[n], b[n])]
    if diffs:
        bad += 1
        n0 = diffs[0][0]
        idx = (a[n0] != b[n0]).nonzero()
        print("trial", trial, "DIFF first:", diffs[0], "all:", [x[0] for x in diffs], "where", idx[:6].tolist(), idx[-3:].tolist(), "shape", list(a[n0].shape))
    else:
        print("trial", trial, "identical")
print("bad trials", bad)

if len(sys.argv) > 5 and bad:
    # structure of the last differing stage3 tensor: per (row // MO) tile for each candidate MO
    a3, b3 = a["dec.stage3"][0], b["dec.stage3"][0]
    rows = (a3 != b3).any(dim=1).nonzero().flatten().tolist()
    print("rows differing", len(rows), "first", rows[:5], "last", rows[-5:])
    for MO in [int(v) for v in sys.argv[5].split(",")]:
        tiles = sorted(set(r // MO for r in rows))
        ctas = sorted(set(t % 148 for t in tiles))
        its = sorted(set(t // 148 for t in tiles))
        print("MO", MO, "tiles", len(tiles), "ctas", ctas[:20], "n_ctas", len(ctas), "iters", its[:40])
    ch = (a3 != b3).any(dim=0).nonzero().flatten().tolist()
    print("channels", ch)
    # runs of consecutive rows
    runs = []
    s0 = p0 = rows[0]
    for r in rows[1:]:
        if r != p0 + 1:
            runs.append((s0, p0 - s0 + 1)); s0 = r
        p0 = r
    runs.append((s0, p0 - s0 + 1))
    print("runs", len(runs), runs[:30])
