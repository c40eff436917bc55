"""CPU oracle of the glue around Synthesizer.infer in the reference's VC pipeline -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product path
(polgen-rvc_b200/) never does.  numpy / torch restatement of rvc/infer/pipeline.py:

    coarse_pitch       get_f0 tail            pipeline.py:186-201
    feature_glue       VC.vc                  pipeline.py:252-270
    change_rms         AudioProcessor         pipeline.py:31-61
    to_int16           VC.pipeline tail       pipeline.py:456-460
    convert_clip       VC.pipeline loop       pipeline.py:329-348, 381-461

Pinned: oracle/make_pipeline_golden.py imports the reference's pipeline.py UNMODIFIED (stub modules
stand in for faiss / torchcrepe / librosa / the F0 predictors, none of which is installed here),
runs VC.get_f0, VC.vc and VC.pipeline with a stub ``net_g`` and writes tests/golden/pipeline/*.npz;
tests/test_pipeline_oracle.py holds this file to those vectors bit for bit.
Third-party arithmetic that is absent here: ``librosa.feature.rms`` (pinned 0.10.2.post1 in the
reference's requirements-win.txt) -- restated below from its published algorithm (center=True,
pad_mode="constant", power = mean(x**2) in float32, sqrt); the golden generator's librosa stub calls
this same function, so change_rms is pinned only up to that restatement.
Scalar promotion: the reference pins numpy 1.23.5, this container has numpy 2.x.  The two differ in one
place on this path -- ``np.abs(audio).max() / 0.99`` is float64 under 1.23.5 and float32 under NEP 50 --
which moves the int16 scale by at most one float32 ulp.  The oracle (and the CUDA kernel) follow what
the reference computes when run HERE (float32), because that is what the golden vectors can pin.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def coarse_pitch(f0: np.ndarray, f0_min: float = 50, f0_max: float = 1100):
    """pipeline.py:150-151,192-201 -> (f0_coarse int, f0bak)"""
    f0_mel_min = 1127 * np.log(1 + f0_min / 700)
    f0_mel_max = 1127 * np.log(1 + f0_max / 700)
    f0 = np.asarray(f0, dtype=np.float64)
    f0bak = f0.copy()
    f0_mel = 1127 * np.log(1 + f0 / 700)
    f0_mel[f0_mel > 0] = (f0_mel[f0_mel > 0] - f0_mel_min) * 254 / (f0_mel_max - f0_mel_min) + 1
    f0_mel[f0_mel <= 1] = 1
    f0_mel[f0_mel > 255] = 255
    return np.rint(f0_mel).astype(int), f0bak


def feature_glue(feats: torch.Tensor, n_audio_frames: int, feats0=None, pitchf=None, protect: float = 0.5):
    """pipeline.py:252-270: feats (1, Th, D) [, feats0 (1, Th, D), pitchf (1, >= p_len)] -> (phone, p_len)"""
    feats = F.interpolate(feats.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)
    do_protect = protect < 0.5 and pitchf is not None and feats0 is not None
    if do_protect:
        feats0 = F.interpolate(feats0.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)
    p_len = n_audio_frames
    if feats.shape[1] < p_len:
        p_len = feats.shape[1]
    if pitchf is not None:
        pitchf = pitchf[:, :p_len]
    feats = feats[:, :p_len]          # the reference leaves feats longer only when 2*Th > p_len (never: HuBERT is shorter)
    if do_protect:
        feats0 = feats0[:, :p_len]
        pitchff = pitchf.clone()
        pitchff[pitchf > 0] = 1
        pitchff[pitchf < 1] = protect
        pitchff = pitchff.unsqueeze(-1)
        feats = feats * pitchff + feats0 * (1 - pitchff)
        feats = feats.to(feats0.dtype)
    return feats, p_len


def librosa_rms(y: np.ndarray, frame_length: int, hop_length: int) -> np.ndarray:
    """librosa.feature.rms(y=y, frame_length=, hop_length=) of librosa 0.10.2 (center=True,
    pad_mode="constant", dtype=float32): (1, 1 + len(y)//hop)"""
    y = np.pad(np.asarray(y), (frame_length // 2, frame_length // 2), mode="constant")
    n_frames = 1 + (y.shape[0] - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n_frames)[None, :]
    x = y[idx]                                            # util.frame: (frame_length, n_frames)
    power = np.mean(np.square(x, dtype=np.float32), axis=-2, keepdims=True)   # util.abs2(x, dtype=float32)
    return np.sqrt(power)


def change_rms(source_audio, source_rate, target_audio, target_rate, rate, rms=librosa_rms):
    """pipeline.py:31-61"""
    rms1 = rms(source_audio, source_rate // 2 * 2, source_rate // 2)
    rms2 = rms(target_audio, target_rate // 2 * 2, target_rate // 2)
    rms1 = F.interpolate(torch.from_numpy(rms1).float().unsqueeze(0), size=target_audio.shape[0], mode="linear").squeeze()
    rms2 = F.interpolate(torch.from_numpy(rms2).float().unsqueeze(0), size=target_audio.shape[0], mode="linear").squeeze()
    rms2 = torch.maximum(rms2, torch.zeros_like(rms2) + 1e-6)
    return target_audio * (torch.pow(rms1, 1 - rate) * torch.pow(rms2, rate - 1)).numpy()


def to_int16(audio_opt: np.ndarray) -> np.ndarray:
    """pipeline.py:456-460 (numpy-2 / NEP 50 scalar semantics: everything stays float32)"""
    audio_opt = np.asarray(audio_opt, dtype=np.float32)
    audio_max = np.float32(np.abs(audio_opt).max()) / np.float32(0.99)
    max_int16 = np.float32(32768)
    if audio_max > 1:
        max_int16 = np.float32(max_int16 / audio_max)
    return (audio_opt * max_int16).astype(np.int16)


def convert_clip(waves, t_pad_tgt: int, source_audio=None, tgt_sr: int = 48000, volume_envelope: float = 1.0):
    """pipeline.py:397,449-460: per-segment waveforms (as returned by VC.vc, untrimmed) -> int16 clip"""
    audio_opt = np.concatenate([np.asarray(w, dtype=np.float32)[t_pad_tgt:-t_pad_tgt] for w in waves])
    if volume_envelope != 1:
        audio_opt = change_rms(source_audio, 16000, audio_opt, tgt_sr, volume_envelope)
    return to_int16(audio_opt), audio_opt
