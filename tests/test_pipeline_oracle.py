"""The pipeline-glue oracle (oracle/pipeline_oracle.py) and the host-side segmentation
(polgen-rvc_b200/segments.py) against golden vectors produced by the LIVE reference pipeline
(oracle/make_pipeline_golden.py: rvc/infer/pipeline.py imported unmodified, stub net_g).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import pipeline_oracle as po
from oracle.make_pipeline_golden import StubHubert, TGT_SR, UPP, clip_audio, clip_f0, stub_waveform

GOLD = os.path.join(ROOT, "tests", "golden", "pipeline")


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_coarse_pitch_matches_reference_get_f0():
    g = gold("coarse_pitch")
    f0 = clip_f0(7, 3000)
    coarse, bak = po.coarse_pitch(f0)
    assert np.array_equal(coarse, g["coarse"]) and np.array_equal(bak, g["f0bak"])
    up, _ = po.coarse_pitch(f0 * pow(2, 5 / 12))            # pipeline.py:183 pitch shift
    assert np.array_equal(up, g["coarse_up5"])
    assert coarse.min() == 1 and coarse.max() == 255       # both clamps exercised


@pytest.mark.parametrize("name", ["vc_v2_protect", "vc_v1_noprotect"])
def test_feature_glue_matches_reference_vc(name):
    g = gold(name)
    feats = torch.from_numpy(g["feats_mixed"])[None]
    feats0 = torch.from_numpy(g["feats0"])[None] if g["feats0"].shape[0] else None
    pitchf = torch.from_numpy(g["pitchf"])[None]
    n_frames = int(g["n_audio"]) // 160
    phone, p_len = po.feature_glue(feats, n_frames, feats0, pitchf, float(g["protect"]))
    assert p_len == int(g["p_len"])
    assert np.array_equal(phone[0].numpy(), g["phone"][:p_len])
    assert np.array_equal(pitchf[0, :p_len].numpy(), g["pitchf_out"])


def replay_pipeline(seconds, seed):
    """the inputs VC.pipeline hands to net_g.infer for the synthetic clip, rebuilt WITHOUT the reference:
    high-pass (pipeline.py:328), our segmentation, the stub HuBERT, oracle feature glue / coarse pitch"""
    from scipy import signal
    from polgen_rvc_b200 import segments as seg
    bh, ah = signal.butter(N=5, Wn=48, btype="high", fs=16000)
    audio = signal.filtfilt(bh, ah, clip_audio(seed, seconds))
    plan = seg.SegmentPlan()
    cuts = seg.split_points(audio, plan)
    frames = seg.segment_frames(audio.shape[0], cuts, plan)
    audio_pad = np.pad(audio, (plan.t_pad, plan.t_pad), mode="reflect")
    p_len_all = audio_pad.shape[0] // 160
    f0 = clip_f0(seed, p_len_all + 8)
    coarse, f0bak = po.coarse_pitch(f0)
    coarse, f0bak = coarse[:p_len_all], f0bak[:p_len_all].astype(np.float32)
    hub = StubHubert(768, 33)
    calls = []
    bounds = [0] + [c // 160 * 160 for c in cuts]
    for i, (start, n) in enumerate(frames):
        s = bounds[i]
        chunk = audio_pad[s: bounds[i + 1] + plan.t_pad2 + 160] if i + 1 < len(bounds) else audio_pad[s:]
        feats = hub.extract_features(torch.from_numpy(chunk).float().view(1, -1), None, 12)[0]
        pitch = torch.from_numpy(coarse[start:start + n])[None]
        pitchf = torch.from_numpy(f0bak[start:start + n])[None]
        phone, p_len = po.feature_glue(feats, chunk.shape[0] // 160, None, pitchf, 0.5)
        calls.append((phone[0].numpy(), pitch[0, :p_len].numpy(), pitchf[0, :p_len].numpy()))
    return audio, calls


@pytest.mark.parametrize("name", ["pipeline_45s", "pipeline_85s_rms"])
def test_whole_pipeline_replay_matches_reference(name):
    g = gold(name)
    audio, calls = replay_pipeline(int(g["seconds"]), int(g["seed"]))
    assert [c[0].shape[0] for c in calls] == list(g["seg_frames"])
    assert [int(c[1][0]) for c in calls] == list(g["seg_first_pitch"])
    waves = [stub_waveform(ph, pi, pf, UPP) for ph, pi, pf in calls]
    pcm, _ = po.convert_clip(waves, TGT_SR * 1, audio, TGT_SR, float(g["vol"]))
    assert pcm.dtype == np.int16 and pcm.shape == g["pcm"].shape
    assert np.array_equal(pcm, g["pcm"])
