"""diagnostic: dense batches, latents vs oracle"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import polgen_rvc_b200 as pg
from oracle import rvc_oracle as orc
cfg = pg.CONFIGS["v2-48k"]
sd = pg.synth_weights(cfg, seed=0, post_std=0.02)
d = torch.device("cuda:0")
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
eng = pg.Engine(cfg, pg.fold_state_dict(sd), 0, flags)
for B, T in ((2, 300), (4, 300), (8, 300), (8, 1000), (3, 1000), (5, 640)):
    inputs = pg.synth_inputs(cfg, B, T, seed=9)
    noise = pg.synth_noise(cfg, B, T, seed=9)
    ref = orc.infer(sd, cfg, *inputs, *noise)
    wave, aux = eng.infer(*[t.to(d) for t in inputs], noise[0].transpose(1, 2).contiguous().to(d),
                          noise[1].reshape(B, -1).contiguous().to(d), 0)
    torch.cuda.synchronize()
    m = aux[2].cpu().transpose(1, 2)
    r = ref[2][2]
    err = (m - r).abs().amax(dim=1)          # [B][T]
    print(f"B={B} T={T} m_p max err {float(err.max()):.4f}; per-row {[round(float(v),4) for v in err.amax(dim=1)]}")
    if float(err.max()) > 1e-2:
        # does our row b match another oracle row?
        for b in range(B):
            best = min(range(B), key=lambda k: float((m[b] - r[k]).abs().max()))
            print("   row", b, "closest oracle row", best, "err", round(float((m[b] - r[best]).abs().max()), 5),
                  "| first t with err>1e-2:", int((err[b] > 1e-2).nonzero()[0]) if (err[b] > 1e-2).any() else None,
                  "ours[0,:3]", [round(float(v), 3) for v in m[b, 0, :3]], "ref", [round(float(v), 3) for v in r[b, 0, :3]])
