"""GPU tests of the pipeline-glue kernels (pg_coarse_pitch, pg_prepare_features, pg_postprocess) and the
batched clip conversion (ClipConverter) against the oracle restatement of rvc/infer/pipeline.py and the
golden vectors of the live reference pipeline.  Integer outputs must be bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from test_gpu_parity import _dev

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden", "pipeline")


def _eng(name="v2-48k"):
    import polgen_rvc_b200 as pg
    cfg = pg.CONFIGS[name]
    return pg.Engine(cfg, pg.fold_state_dict(pg.synth_weights(cfg, seed=0)), 0)


def test_coarse_pitch_bit_exact():
    import polgen_rvc_b200 as pg
    from oracle import pipeline_oracle as po
    from oracle.make_pipeline_golden import clip_f0
    g = np.load(os.path.join(GOLD, "coarse_pitch.npz"))
    eng = _eng()
    d = _dev()
    f0 = clip_f0(7, 3000)
    pitch, pitchf = pg.coarse_pitch(eng, torch.from_numpy(f0).to(d))
    assert np.array_equal(pitch.cpu().numpy(), g["coarse"])                       # live reference get_f0
    assert np.array_equal(pitchf.cpu().numpy(), g["f0bak"].astype(np.float32))
    big = np.random.default_rng(0).uniform(0, 1300, 200000)
    big[::7] = 0
    pitch, _ = pg.coarse_pitch(eng, torch.from_numpy(big).to(d))
    assert np.array_equal(pitch.cpu().numpy(), po.coarse_pitch(big)[0])           # oracle, 2e5 frames
    f32 = big.astype(np.float32)
    pitch, _ = pg.coarse_pitch(eng, torch.from_numpy(f32).to(d))
    assert np.array_equal(pitch.cpu().numpy(), po.coarse_pitch(f32.astype(np.float64))[0])


@pytest.mark.parametrize("name", ["vc_v2_protect", "vc_v1_noprotect"])
def test_prepare_features_bit_exact(name):
    import polgen_rvc_b200 as pg
    g = np.load(os.path.join(GOLD, name + ".npz"))
    eng = _eng("v2-48k" if "v2" in name else "v1-40k")
    d = _dev()
    feats = torch.from_numpy(g["feats_mixed"]).to(d)
    feats0 = torch.from_numpy(g["feats0"]).to(d) if g["feats0"].shape[0] else None
    pitchf = torch.from_numpy(g["pitchf"]).to(d)
    phone = pg.prepare_features(eng, feats, int(g["n_audio"]) // 160, feats0, pitchf, float(g["protect"]))
    p_len = int(g["p_len"])
    assert phone.shape == (p_len, feats.shape[1])
    assert np.array_equal(phone.cpu().numpy(), g["phone"][:p_len])               # live reference VC.vc


def test_postprocess_int16_bit_exact_and_change_rms():
    import polgen_rvc_b200 as pg
    from oracle import pipeline_oracle as po
    eng = _eng()
    d = _dev()
    rng = np.random.default_rng(5)
    for n, amp in ((48000 * 3 + 17, 1.7), (48000 * 2, 0.4), (1000, 0.99), (48000 * 61, 1.2)):
        x = (rng.standard_normal(n) * amp / 3).astype(np.float32)
        pcm = pg.postprocess(eng, torch.from_numpy(x).to(d))
        torch.cuda.synchronize()
        assert np.array_equal(pcm.cpu().numpy(), po.to_int16(x)), (n, amp)
    # change_rms path: f32 arithmetic (different summation order than numpy): relative 1e-5, int16 +-1
    n_src, n = 16000 * 7 + 123, 48000 * 7 + 400
    src = (rng.standard_normal(n_src) * (0.2 + np.abs(np.sin(np.arange(n_src) / 9000.0)))).astype(np.float64)
    x = (rng.standard_normal(n) * 0.3).astype(np.float32)
    want_f = po.change_rms(src, 16000, x, 48000, 0.25).astype(np.float32)
    pinned = torch.empty(n, dtype=torch.int16).pin_memory()
    pcm, fl = pg.postprocess(eng, torch.from_numpy(x).to(d), torch.from_numpy(src).to(d), 16000, 48000, 0.25,
                             pcm_out=pinned, want_float=True)
    torch.cuda.synchronize()
    got_f = fl.cpu().numpy()
    assert np.abs(got_f - want_f).max() <= 2e-5 * np.abs(want_f).max()
    assert np.array_equal(pcm.numpy(), po.to_int16(got_f))                        # exact given the same f32 audio
    assert np.abs(pcm.numpy().astype(np.int32) - po.to_int16(want_f).astype(np.int32)).max() <= 1


def test_pipeline_replay_on_device_matches_reference_pcm():
    """The golden VC.pipeline runs (live reference, stub net_g): their waveforms, replayed through the
    device trim/concat + pg_postprocess, give the reference's int16 clip bit for bit (vol = 1) / +-1 (rms)."""
    import polgen_rvc_b200 as pg
    from oracle.make_pipeline_golden import TGT_SR, UPP, stub_waveform
    from test_pipeline_oracle import replay_pipeline
    eng = _eng()
    d = _dev()
    for name in ("pipeline_45s", "pipeline_85s_rms"):
        g = np.load(os.path.join(GOLD, name + ".npz"))
        audio, calls = replay_pipeline(int(g["seconds"]), int(g["seed"]))
        waves = [torch.from_numpy(stub_waveform(ph, pi, pf, UPP)).to(d) for ph, pi, pf in calls]
        clip = torch.cat([w[TGT_SR:-TGT_SR] for w in waves])
        vol = float(g["vol"])
        pcm = pg.postprocess(eng, clip, torch.from_numpy(audio.copy()).to(d) if vol != 1 else None, 16000, TGT_SR, vol)
        torch.cuda.synchronize()
        diff = np.abs(pcm.cpu().numpy().astype(np.int32) - g["pcm"].astype(np.int32)).max()
        assert diff == 0 if vol == 1 else diff <= 1, (name, diff)


def test_clip_converter_equals_segment_loop():
    """ClipConverter (ragged batches over two lanes -> one device buffer -> int16) against the reference
    loop restated with our own B=1 calls: per-segment pg_infer_segments, host trim/concat, oracle int16."""
    import polgen_rvc_b200 as pg
    from oracle import pipeline_oracle as po
    cfg = pg.CONFIGS["v2-40k"]
    folded = pg.fold_state_dict(pg.synth_weights(cfg, seed=0))
    sched = pg.SegmentScheduler(cfg, folded, 0, lanes=2, max_batch=2)
    conv = pg.ClipConverter(sched, cfg.sr)
    d = _dev()
    frames = [330, 250, 410]
    segs = [pg.synth_inputs(cfg, 1, T, seed=80 + i) for i, T in enumerate(frames)]
    dev_segs = [[t.to(d) if j != 4 else t for j, t in enumerate(s)] for s in segs]
    src = torch.from_numpy(np.random.default_rng(1).standard_normal(16000 * 8) * 0.2)
    for vol in (1.0, 0.5):
        pcm = conv.convert(dev_segs, source_audio=src, volume_envelope=vol, seed=11).clone()
        # the same batches, one engine, then the host-side tail of the reference
        eng = pg.Engine(cfg, folded, 0)
        waves = [None] * len(frames)
        for k, idx in enumerate(sched.last_batches):
            ws, _ = eng.infer_segments([sched._segment_dict(dev_segs[i]) for i in idx], seed=11 + k)
            torch.cuda.synchronize()
            for i, w in zip(idx, ws):
                waves[i] = w.cpu().numpy()
        want, _ = po.convert_clip(waves, cfg.sr * 1, src.numpy(), cfg.sr, vol)
        assert pcm.shape == want.shape and pcm.dtype == torch.int16
        diff = np.abs(pcm.numpy().astype(np.int32) - want.astype(np.int32)).max()
        assert diff == 0 if vol == 1.0 else diff <= 1, (vol, diff)
