"""Role-level timeline of the TextEncoder / flow GEMM kernel (conv_umma_kernel, CTA 0 of every launch):
    PG_UMMA_DEBUG=8 python tools/umma_trace.py 2> gpurun_out/umma_trace.txt
One flow_reverse + one text_encoder call on the bench clip's ragged batch; the kernel prints
`TR <ns since first event> role ev tile` lines (roles: 0 weight TMA, 1 MMA, 2 loader, 3 epilogue, 4 CTA)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import polgen_rvc_b200 as pg  # noqa: E402
from polgen_rvc_b200 import _lib  # noqa: E402

cfg = pg.CONFIGS["v2-48k"]
eng = pg.Engine(cfg, pg.fold_state_dict(pg.synth_weights(cfg, seed=0)), 0, _lib.PG_FLAG_NO_GRAPHS)
d = torch.device("cuda:0")
lens = [3435, 2965]
phone, lengths, pitch, f0, sid = pg.synth_inputs(cfg, 2, max(lens), seed=0)
lengths = torch.tensor(lens)
m, l = eng.text_encoder(phone.to(d), lengths.to(d), pitch.to(d))
torch.cuda.synchronize()
