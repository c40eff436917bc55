"""A/B check of an env-selected kernel variant at bench-like size: run once without the env switch
(saves the waveform), once with it (compares).   python tools/swap_check.py save|cmp [T]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import polgen_rvc_b200 as pg  # noqa: E402

mode = sys.argv[1]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1400
cfg = pg.CONFIGS["v2-48k"]
eng = pg.Engine(cfg, pg.fold_state_dict(pg.synth_weights(cfg, seed=0)), 0)
d = torch.device("cuda:0")
inp = [t.to(d) for t in pg.synth_inputs(cfg, 1, T, seed=1)]
wave, _ = eng.infer(*inp, None, None, 5, want_aux=False)
torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", f"swap_ref_{T}.pt")
if mode == "save":
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(wave.cpu(), path)
    print("saved", tuple(wave.shape), float(wave.abs().max()))
else:
    ref = torch.load(path).to(d)
    err = (wave - ref).abs().max().item()
    snr = 10 * torch.log10((ref ** 2).sum() / ((wave - ref) ** 2).sum().clamp_min(1e-30)).item()
    print(f"cmp T={T}: max-abs {err:.3e}  SNR {snr:.1f} dB  finite={bool(torch.isfinite(wave).all())}")
