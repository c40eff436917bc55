"""One ragged decode of the bench clip ([3435, 2965] frames, v2-48k) per iteration -- the command ncu wraps:
    ncu --metrics ... -k regex:"pair_planes|conv_planes" --launch-skip <n> --launch-count <n> python tools/ncu_decode.py [iters] [flags]
Direct launches (PG_FLAG_NO_GRAPHS) so every kernel is a separate ncu launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import polgen_rvc_b200 as pg  # noqa: E402
from polgen_rvc_b200 import _lib  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
flags = _lib.PG_FLAG_NO_GRAPHS | (int(sys.argv[2]) if len(sys.argv) > 2 else 0)
cfg = pg.CONFIGS["v2-48k"]
eng = pg.Engine(cfg, pg.fold_state_dict(pg.synth_weights(cfg, seed=0)), 0, flags)
d = torch.device("cuda:0")
rows = []
for i, T in enumerate((3435, 2965)):
    phone, _, pitch, f0, _ = pg.synth_inputs(cfg, 1, T, seed=100 + i)
    rows.append({"phone": phone[0].to(d), "pitch": pitch[0].to(d), "f0": f0[0].to(d), "sid": 0})
for it in range(iters):
    waves, _ = eng.infer_segments(rows, seed=it)
    torch.cuda.synchronize()
print("launches per decode:", eng.launch_count())
