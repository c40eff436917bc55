import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_files():
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def load_case(path):
    """-> (cfg, state_dict, inputs, noise, golden) for one fixture, everything regenerated from seeds."""
    import numpy as np
    import polgen_rvc_b200 as pg
    name = os.path.basename(path).split("_")[0]
    g = np.load(path)
    B, T, seed = (int(v) for v in g["meta"])
    cfg = pg.CONFIGS[name]
    sd = pg.synth_weights(cfg, seed=seed)
    inputs = pg.synth_inputs(cfg, B, T, seed=seed)
    noise = pg.synth_noise(cfg, B, T, seed=seed)
    return cfg, sd, inputs, noise, g


def snr_db(got, want):
    import torch
    got = torch.as_tensor(got).double().flatten()
    want = torch.as_tensor(want).double().flatten()
    return float(10 * torch.log10((want ** 2).sum() / ((got - want) ** 2).sum().clamp_min(1e-300)))


@pytest.fixture(scope="session")
def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "rvc", "lib", "algorithm"))
