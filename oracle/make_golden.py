"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

Imports ``rvc.lib.algorithm.synthesizers.Synthesizer`` unmodified from
/root/reference, builds it exactly like ``rvc/infer/infer.py:92-101``
(``Synthesizer(*cfg, use_f0=1, input_dim=..., is_half=False)``, ``del enc_q``,
``load_state_dict(strict=False)``, ``.eval().float()``), loads the deterministic
legacy-spelled weights from ``configs.synth_weights`` and replays the two
normal draws (synthesizers.py:174, generators.py:154) from
``configs.synth_noise`` by temporarily wrapping ``torch.randn_like``.

The fixtures hold only OUTPUTS (waveform + a few intermediate taps); weights,
inputs and noise are regenerated from seeds by the tests, so nothing here
needs /root/reference at test time.

    python oracle/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import polgen_rvc_b200 as pg  # noqa: E402

CASES = [  # (config, batch, frames, seed)
    ("v2-40k", 1, 24, 0),
    ("v2-48k", 1, 16, 1),
    ("v2-32k", 2, 12, 2),
    ("v1-40k", 1, 13, 3),
]


def run_reference(cfg, sd, phone, lengths, pitch, f0, sid, eps_zp, eps_src):
    from rvc.lib.algorithm.synthesizers import Synthesizer
    net = Synthesizer(*cfg.ctor_args(), use_f0=1, input_dim=cfg.input_dim, is_half=False)
    del net.enc_q
    missing = net.load_state_dict(sd, strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys, missing
    net.eval().float()
    taps = {}
    hooks = [
        net.dec.m_source.register_forward_hook(lambda m, i, o: taps.__setitem__("source", o[0].detach().clone())),
        net.dec.conv_pre.register_forward_hook(lambda m, i, o: taps.__setitem__("dec.conv_pre_nocond", o.detach().clone())),
    ]
    queue = [eps_zp, eps_src]
    real = torch.randn_like

    def fake(t, *a, **k):
        e = queue.pop(0)
        assert e.shape == t.shape, (e.shape, t.shape)
        return e.to(t.dtype)

    torch.randn_like = fake
    try:
        with torch.no_grad():
            o, x_mask, (z, z_p, m_p, logs_p) = net.infer(phone, lengths, pitch, f0, sid)
    finally:
        torch.randn_like = real
        for h in hooks:
            h.remove()
    assert not queue
    return dict(o=o, x_mask=x_mask, z=z, z_p=z_p, m_p=m_p, logs_p=logs_p, source=taps["source"])


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, B, T, seed in CASES:
        cfg = pg.CONFIGS[name]
        sd = pg.synth_weights(cfg, seed=seed)
        phone, lengths, pitch, f0, sid = pg.synth_inputs(cfg, B, T, seed=seed)
        eps_zp, eps_src = pg.synth_noise(cfg, B, T, seed=seed)
        res = run_reference(cfg, sd, phone, lengths, pitch, f0, sid, eps_zp, eps_src)
        # also a noise-free source for the <=1e-5 sine check
        res0 = run_reference(cfg, sd, phone, lengths, pitch, f0, sid, eps_zp, torch.zeros_like(eps_src))
        path = os.path.join(out_dir, f"{name}_B{B}_T{T}_s{seed}.npz")
        np.savez_compressed(
            path,
            o=res["o"].numpy().astype(np.float32),
            z=res["z"].numpy().astype(np.float32),
            z_p=res["z_p"].numpy().astype(np.float32),
            m_p=res["m_p"].numpy().astype(np.float32),
            logs_p=res["logs_p"].numpy().astype(np.float32),
            source=res["source"].numpy().astype(np.float32),
            source_nonoise=res0["source"].numpy().astype(np.float32),
            meta=np.array([B, T, seed]),
        )
        print(path, os.path.getsize(path), "bytes", "o.std", float(res["o"].std()))


if __name__ == "__main__":
    main()
