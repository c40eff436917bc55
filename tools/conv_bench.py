"""Single-shape conv micro-benchmark through pg_op_conv1d_f16 (for ncu captures).
    python tools/conv_bench.py C K dil L B [iters] [impl]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from polgen_rvc_b200 import _lib  # noqa: E402

Cc, K, dil, L, B = (int(v) for v in sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
impl = int(sys.argv[7]) if len(sys.argv) > 7 else 1
lib = _lib.load()
d = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randn(B, L, Cc, generator=g).half().to(d)
res = torch.randn(B, L, Cc, generator=g).half().to(d)
w = (torch.randn(Cc, Cc, K, generator=g) / (Cc * K) ** 0.5).contiguous()
bias = torch.randn(Cc, generator=g) * 0.1
y = torch.zeros(B, L, Cc, device=d, dtype=torch.half)
ms = C.c_float(0)
for _ in range(2):
    rc = lib.pg_op_conv1d_f16(0, impl, B, L, Cc, Cc, K, dil, C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()),
                              C.c_void_p(bias.data_ptr()), C.c_float(0.1), C.c_float(1.0), C.c_void_p(res.data_ptr()),
                              C.c_void_p(y.data_ptr()), iters, C.byref(ms))
    assert rc == 0, lib.pg_last_error()
fl = 2.0 * B * L * Cc * Cc * K
print(f"C={Cc} K={K} dil={dil} L={L} B={B}: {ms.value / iters * 1e3:.1f} us/launch, {fl / (ms.value / iters * 1e-3) / 1e12:.1f} TFLOP/s")
