"""Host-side logic on CPU: drop-in class surface, weight-norm fold, config tables, segment
sharding arithmetic."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REFERENCE

import polgen_rvc_b200 as pg
from polgen_rvc_b200 import segments as seg
from oracle import rvc_oracle as orc


@pytest.mark.parametrize("name", list(pg.CONFIGS))
def test_state_dict_inventory_and_legacy_load(name):
    cfg = pg.CONFIGS[name]
    net = pg.Synthesizer(*cfg.ctor_args(), use_f0=1, input_dim=cfg.input_dim, is_half=False)
    del net.enc_q                                   # infer.py:99
    keys = set(net.state_dict().keys())
    legacy = pg.synth_weights(cfg, seed=1)
    assert {k for k, _, _ in pg.configs.state_dict_shapes(cfg)} == set(legacy)
    res = net.load_state_dict(legacy, strict=False)  # infer.py:100, legacy weight_g/weight_v spelling
    assert not res.missing_keys and not res.unexpected_keys
    assert len(keys) == 457
    folded = pg.fold_state_dict(net.state_dict())
    want = orc.fold_weight_norm(legacy)
    assert set(folded) == set(want)
    for k in want:
        assert torch.allclose(folded[k], want[k], rtol=1e-6, atol=1e-8), k
    net.eval().float()                              # infer.py:101-102 surface


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="live reference not present")
@pytest.mark.parametrize("name", ["v2-48k", "v1-40k"])
def test_state_dict_keys_equal_reference(name):
    sys.path.insert(0, REFERENCE)
    from rvc.lib.algorithm.synthesizers import Synthesizer as Ref
    cfg = pg.CONFIGS[name]
    ref = Ref(*cfg.ctor_args(), use_f0=1, input_dim=cfg.input_dim, is_half=False)
    del ref.enc_q
    mine = pg.Synthesizer(*cfg.ctor_args(), use_f0=1, input_dim=cfg.input_dim, is_half=False)
    del mine.enc_q
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a) == list(b)
    assert all(a[k].shape == b[k].shape for k in a)
    assert not mine.load_state_dict(a, strict=True).missing_keys


def test_constructor_contract():
    cfg = pg.CONFIGS["v2-40k"]
    with pytest.raises(KeyError):
        pg.Synthesizer(*cfg.ctor_args(), use_f0=1, input_dim=768)            # is_half is required
    with pytest.raises(NotImplementedError):
        pg.Synthesizer(*cfg.ctor_args(), use_f0=0, input_dim=768, is_half=False)
    a = pg.SynthesizerTrnMs768NSFsid(*cfg.ctor_args(), is_half=False)
    b = pg.SynthesizerTrnMs256NSFsid(*cfg.ctor_args(), is_half=False)
    assert a.enc_p.emb_phone.weight.shape == (192, 768)
    assert b.enc_p.emb_phone.weight.shape == (192, 256)
    assert a.flow.flows[0].post.weight.abs().sum() == 0     # residuals.py:207-208


def test_generator_flops_match_survey():
    want = {"v2-48k": 110.15, "v2-40k": 91.45, "v2-32k": 76.89, "v1-40k": 91.45}
    for name, gf in want.items():
        got = pg.CONFIGS[name].generator_flops_per_frame() * 100 / 1e9
        assert abs(got - gf) < 0.01, (name, got)


def test_synth_inputs_follow_pipeline_quantisation():
    cfg = pg.CONFIGS["v2-48k"]
    phone, lengths, pitch, f0, sid = pg.synth_inputs(cfg, 2, 400, seed=0)
    assert phone.shape == (2, 400, 768) and pitch.dtype == torch.long
    assert int(pitch.min()) >= 1 and int(pitch.max()) <= 255
    assert (f0 == 0).any() and (f0 > 0).any()
    assert bool((pitch[f0 == 0] == 1).all())                # unvoiced -> coarse 1 (pipeline.py:198)
    assert float(f0.max()) <= 800.0 + 1e-3 and float(f0[f0 > 0].min()) >= 80.0 - 1e-3


def _reference_split(audio, plan):
    # literal restatement of pipeline.py:330-344 for the test
    window = seg.WINDOW
    audio_pad = np.pad(audio, (window // 2, window // 2), mode="reflect")
    opt_ts = []
    if audio_pad.shape[0] > plan.t_max:
        audio_sum = np.zeros_like(audio)
        for i in range(window):
            audio_sum += audio_pad[i:i - window]
        for t in range(plan.t_center, audio.shape[0], plan.t_center):
            win = np.abs(audio_sum[t - plan.t_query:t + plan.t_query])
            opt_ts.append(t - plan.t_query + np.where(win == win.min())[0][0])
    return opt_ts


def test_split_points_match_reference_loop():
    rng = np.random.default_rng(0)
    plan = seg.SegmentPlan()
    for seconds in (10, 60, 95):
        n = 16000 * seconds
        audio = rng.standard_normal(n) * (0.2 + np.abs(np.sin(np.arange(n) / 16000.0 * 0.7)))
        assert seg.split_points(audio, plan) == [int(v) for v in _reference_split(audio, plan)]
    assert seg.split_points(rng.standard_normal(16000 * 10), plan) == []


def test_segment_frames_cover_clip_with_overlap():
    plan = seg.SegmentPlan()
    n = 16000 * 60
    cuts = [16000 * 37 + 123]
    fr = seg.segment_frames(n, cuts, plan)
    assert len(fr) == 2
    p_len = (n + 2 * plan.t_pad) // seg.WINDOW
    assert fr[0][0] == 0 and fr[-1][0] + fr[-1][1] == p_len
    # consecutive segments overlap by 2*x_pad seconds of frames (trimmed later)
    assert fr[0][0] + fr[0][1] - fr[1][0] == 2 * plan.x_pad * 100
    total_out = sum(c - 2 * plan.x_pad * 100 for _, c in fr)
    assert total_out == p_len - 2 * plan.x_pad * 100


def test_trim_and_concat():
    plan = seg.SegmentPlan()
    sr = 48000
    waves = [np.arange(sr * 3, dtype=np.float32), np.arange(sr * 4, dtype=np.float32)]
    out = seg.trim_and_concat(waves, sr, plan)
    assert out.shape[0] == sr * 1 + sr * 2
    assert out[0] == sr and out[sr] == sr


def test_plan_shards_balanced_and_complete():
    lengths = [1000] * 512
    for world in (1, 2, 4, 8):
        bins = seg.plan_shards(lengths, world)
        assert sorted(i for b in bins for i in b) == list(range(512))
        assert all(len(b) == 512 // world for b in bins)
    bins = seg.plan_shards([4000, 2400], 8)
    assert [len(b) for b in bins] == [1, 1, 0, 0, 0, 0, 0, 0]
    bins = seg.plan_shards([5, 9, 3, 7, 7, 1], 2)
    loads = [sum([5, 9, 3, 7, 7, 1][i] for i in b) for b in bins]
    assert abs(loads[0] - loads[1]) <= 3 and sum(loads) == 32
    batches = seg.batch_equal_lengths([0, 1, 2, 3, 4], [10, 20, 10, 10, 20], max_batch=2)
    assert sorted(map(tuple, batches)) == [(0, 2), (1, 4), (3,)]


def test_bench_reference_arm_line_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the
    same metric / unit / workload as our arm, no GPU work; kind "reference" when build() vendored the
    unmodified reference module under baseline/_ref (this container), "port" (oracle) otherwise."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    have_ref = os.path.isfile(os.path.join(root, "baseline", "_ref", "rvc", "lib", "algorithm", "synthesizers.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "[3435, 2965]" in d["config"]["workload"] and "sample" in d["config"]
    assert d["gpu_launches"] == 0


def test_plan_shards_properties_random():
    """Randomised properties of the segment sharding (greedy longest-first): every segment on exactly
    one rank, deterministic, and the classic LPT bound max_load <= mean_load + max_item."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(min_value=1, max_value=6000), min_size=0, max_size=64),
           st.integers(min_value=1, max_value=8))
    def check(lengths, world):
        bins = seg.plan_shards(lengths, world)
        assert len(bins) == world
        assert sorted(i for b in bins for i in b) == list(range(len(lengths)))
        assert bins == seg.plan_shards(list(lengths), world)
        assert all(b == sorted(b) for b in bins)
        if lengths:
            loads = [sum(lengths[i] for i in b) for b in bins]
            assert max(loads) <= sum(lengths) / world + max(lengths)
    check()


def test_scheduler_batch_planning_without_engine():
    """SegmentScheduler.plan_batches (host logic only): longest first, bounded rows and padded frames."""
    import polgen_rvc_b200 as pg

    class FakeEngine:
        def padded_frames(self, T):
            q = 32 if T <= 512 else (64 if T <= 2048 else 128)
            return (T + q - 1) // q * q

    s = pg.SegmentScheduler.__new__(pg.SegmentScheduler)
    s.engines, s.max_batch, s.max_batch_frames = [FakeEngine()], 3, 8000
    lengths = [3435, 2965, 100, 4100, 1000, 1000, 1000, 50]
    batches = s.plan_batches(lengths)
    assert sorted(i for b in batches for i in b) == list(range(len(lengths)))
    for b in batches:
        top = max(FakeEngine().padded_frames(lengths[i]) for i in b)
        assert len(b) <= 3 and (len(b) == 1 or len(b) * top <= 8000)
    assert batches[0][0] == 3            # longest first
    assert s.plan_batches([]) == []


def test_time_tile_planning():
    """plan_tiles: tiles cover [0, T) exactly once, interior cuts aligned, never more tiles than asked."""
    import polgen_rvc_b200 as pg
    for T, n in ((3435, 8), (3435, 2), (100, 8), (7, 4), (1, 3), (4100, 5)):
        tiles = pg.plan_tiles(T, n)
        assert 1 <= len(tiles) <= n
        assert tiles[0][0] == 0 and tiles[-1][1] == T
        assert all(tiles[i][1] == tiles[i + 1][0] for i in range(len(tiles) - 1))
        assert all(b > a for a, b in tiles)
        assert all(a % 8 == 0 for a, _ in tiles)
