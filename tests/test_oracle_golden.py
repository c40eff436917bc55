"""The oracle (oracle/rvc_oracle.py) against the golden vectors generated from the LIVE
reference (oracle/make_golden.py), and against the live reference itself when it is present.
CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REFERENCE, golden_files, load_case, snr_db
from oracle import rvc_oracle as orc


@pytest.mark.parametrize("path", golden_files(), ids=os.path.basename)
def test_oracle_matches_golden(path):
    cfg, sd, (phone, lengths, pitch, f0, sid), (eps_zp, eps_src), g = load_case(path)
    taps = {}
    o, mask, (z, z_p, m_p, logs_p) = orc.infer(sd, cfg, phone, lengths, pitch, f0, sid, eps_zp, eps_src,
                                                taps=taps)
    assert o.shape == g["o"].shape
    assert snr_db(o, g["o"]) > 100.0
    assert np.abs(o.numpy() - g["o"]).max() < 1e-5
    for name, got in (("z", z), ("z_p", z_p), ("m_p", m_p), ("logs_p", logs_p)):
        assert np.abs(got.numpy() - g[name]).max() < 1e-4, name
    assert np.abs(taps["source"].numpy() - g["source"]).max() < 1e-6


@pytest.mark.parametrize("path", golden_files(), ids=os.path.basename)
def test_exact_sine_within_1e5_of_reference(path):
    """The closed-form fp64 phase (what the CUDA kernel evaluates) vs the reference's source."""
    cfg, sd, (phone, lengths, pitch, f0, sid), _, g = load_case(path)
    W = orc.fold_weight_norm(sd)
    src, _ = orc.sine_source(W, cfg, f0, None, exact=True)
    assert np.abs(src.numpy() - g["source_nonoise"]).max() <= 1e-5


def test_exact_sine_long_clip():
    """10 s and 40 s clips: closed form vs the reference op sequence stays <= 1e-5."""
    import polgen_rvc_b200 as pg
    cfg = pg.CONFIGS["v2-48k"]
    W = {"dec.m_source.l_linear.weight": torch.tensor([[1.0]]), "dec.m_source.l_linear.bias": torch.tensor([0.0])}
    for T in (1000, 4000):
        _, _, _, f0, _ = pg.synth_inputs(cfg, 1, T, seed=5)
        a, sa = orc.sine_source(W, cfg, f0, None, exact=False)
        b, sb = orc.sine_source(W, cfg, f0, None, exact=True)
        assert (sa - sb).abs().max().item() <= 1e-5, T


def test_fold_weight_norm_both_spellings():
    import polgen_rvc_b200 as pg
    cfg = pg.CONFIGS["v1-40k"]
    sd = pg.synth_weights(cfg, seed=7)
    new = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            new[k[:-len(".weight_g")] + ".parametrizations.weight.original0"] = v
        elif k.endswith(".weight_v"):
            new[k[:-len(".weight_v")] + ".parametrizations.weight.original1"] = v
        else:
            new[k] = v
    a, b = orc.fold_weight_norm(sd), orc.fold_weight_norm(new)
    assert set(a) == set(b)
    assert all(torch.equal(a[k], b[k]) for k in a)
    # ConvTranspose1d: norm is per INPUT channel (dim 0)
    w = a["dec.ups.0.weight"]
    v, g = sd["dec.ups.0.weight_v"], sd["dec.ups.0.weight_g"]
    assert torch.allclose(w.reshape(w.shape[0], -1).norm(dim=1), g.reshape(-1), rtol=1e-5)
    assert torch.allclose(w[3] / w[3].norm(), v[3] / v[3].norm(), atol=1e-6)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="live reference not present on this box")
def test_oracle_matches_live_reference_masked_batch():
    """Ragged lengths + re-randomised post convs through the unmodified reference module."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.make_golden import run_reference
    import polgen_rvc_b200 as pg
    cfg = pg.CONFIGS["v2-32k"]
    B, T = 3, 14
    sd = pg.synth_weights(cfg, seed=11, post_std=0.05)
    phone, lengths, pitch, f0, sid = pg.synth_inputs(cfg, B, T, seed=11)
    lengths = torch.tensor([14, 9, 3])
    sid = torch.tensor([0, 5, 108])
    eps_zp, eps_src = pg.synth_noise(cfg, B, T, seed=11)
    ref = run_reference(cfg, sd, phone, lengths, pitch, f0, sid, eps_zp, eps_src)
    o, mask, (z, z_p, m_p, logs_p) = orc.infer(sd, cfg, phone, lengths, pitch, f0, sid, eps_zp, eps_src)
    assert snr_db(o, ref["o"]) > 100.0
    assert (z - ref["z"]).abs().max() < 1e-4
    assert torch.equal(mask, ref["x_mask"])
