"""diagnostic: the same ragged call repeated -> how many calls differ from the first?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polgen_rvc_b200 as pg
cfg = pg.CONFIGS[sys.argv[2] if len(sys.argv) > 2 else "v2-40k"]
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 150
folded = pg.fold_state_dict(pg.synth_weights(cfg, seed=0))
d = torch.device("cuda:0")
frames = [int(v) for v in os.environ.get("DIAG_FRAMES", "410,330").split(",")]
rows = []
for i, T in enumerate(frames):
    phone, _, pitch, f0, _ = pg.synth_inputs(cfg, 1, T, seed=80 + i)
    rows.append({"phone": phone[0].to(d), "pitch": pitch[0].to(d), "f0": f0[0].to(d), "sid": 0})
eng = pg.Engine(cfg, folded, 0, flags)
st = torch.cuda.Stream()
ref, bad, worst, sizes = None, 0, 0.0, {}
for rep in range(reps):
    with torch.cuda.stream(st):
        w, _ = eng.infer_segments(rows, seed=11)
    st.synchronize()
    o = torch.cat(w)
    if ref is None:
        ref = o.clone()
        continue
    n = int((o != ref).sum())
    if n:
        idx = (o != ref).nonzero().flatten()
        print("  call", rep, "n", n, "first", int(idx[0]), "last", int(idx[-1]), "maxdiff", float((o - ref).abs().max()))
        bad += 1
        worst = max(worst, float((o - ref).abs().max()))
        sizes[n] = sizes.get(n, 0) + 1
print(f"{reps} calls: {bad} differ from the first, worst {worst:.2e}, sizes {sizes}")
import ctypes as C
lib = eng.lib
try:
    f = C.CDLL(lib._name).pg_debug_pair_counter
    f.restype = C.c_uint64
    vals = [f(i) for i in range(10)]
    import struct
    fl = lambda v: struct.unpack("f", struct.pack("I", v & 0xffffffff))[0]
    print("dbg: compared", vals[0], "mismatch", vals[1], "early", fl(vals[2]), fl(vals[8]), "late", fl(vals[3]), fl(vals[9]), "t", vals[4], "plane", vals[5], "cta", vals[6], "tile#", vals[7])
except Exception as e:
    print("no dbg", e)
