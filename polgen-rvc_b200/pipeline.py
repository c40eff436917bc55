"""The batched, device-resident replacement of the reference's per-clip conversion loop.

The reference's ``VC.pipeline`` (``rvc/infer/pipeline.py:287-467``) walks the silence-split
segments of a clip one at a time: ``VC.vc`` (``:203-286``) interpolates the HuBERT features,
applies the ``protect`` mix, calls ``net_g.infer`` at B = 1, pulls the waveform to the host
(``.data.cpu().float().numpy()``), trims ``t_pad_tgt`` and finally concatenates, applies
``change_rms`` (``:31-61``) and converts to int16 in numpy (``:449-460``).  Everything between
the HuBERT / F0 predictors and the int16 result is done here on the GPU, through the C ABI
(include/polgen_rvc.h):

* ``coarse_pitch``       -- tail of ``VC.get_f0``            (``:186-201``)  pg_coarse_pitch
* ``prepare_features``   -- feature glue of ``VC.vc``        (``:252-270``)  pg_prepare_features
* ``ClipConverter``      -- the segment loop + post step    (``:381-461``)  pg_infer_segments + pg_postprocess

HuBERT, RMVPE+/FCPE/crepe and the faiss retrieval stay in the reference and hand their outputs in.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .segments import SegmentPlan
from .synthesizer import Engine, SegmentScheduler


def _stream(device: int):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def coarse_pitch(eng: Engine, f0: torch.Tensor, f0_min: float = 50.0, f0_max: float = 1100.0):
    """f0 [n] (f64 or f32, Hz, on the engine's GPU) -> (pitch i64 [n] in 1..255, pitchf f32 [n])."""
    if f0.dtype not in (torch.float64, torch.float32) or not f0.is_cuda:
        raise ValueError("f0 must be a CUDA f64 / f32 tensor")
    f0 = f0.contiguous()
    n = f0.numel()
    pitch = torch.empty(n, dtype=torch.int64, device=f0.device)
    pitchf = torch.empty(n, dtype=torch.float32, device=f0.device)
    _lib.check(eng.lib.pg_coarse_pitch(eng._h, _stream(eng.device), n, C.c_void_p(f0.data_ptr()),
                                       _lib.PG_F64 if f0.dtype == torch.float64 else _lib.PG_F32,
                                       float(f0_min), float(f0_max), C.c_void_p(pitch.data_ptr()),
                                       C.c_void_p(pitchf.data_ptr())), "pg_coarse_pitch")
    return pitch, pitchf


def prepare_features(eng: Engine, feats: torch.Tensor, n_audio_frames: int, feats0: Optional[torch.Tensor] = None,
                     pitchf: Optional[torch.Tensor] = None, protect: float = 0.5):
    """feats [Th][D] f32 (HuBERT output, after the optional index mix) -> phone [p_len][D] with
    p_len = min(n_audio_frames, 2*Th): x2 nearest interpolate + protect mix (needs feats0, pitchf)."""
    feats = feats.reshape(-1, feats.shape[-1]).contiguous()
    Th, D = feats.shape
    if D != eng.cfg.input_dim:
        raise ValueError(f"feature width {D} != input_dim {eng.cfg.input_dim}")
    p_len = min(int(n_audio_frames), 2 * Th)
    if feats0 is not None:
        feats0 = feats0.reshape(-1, D).contiguous()
    if pitchf is not None:
        pitchf = pitchf.reshape(-1)[:p_len].contiguous()
    out = torch.empty(p_len, D, dtype=torch.float32, device=feats.device)
    got = C.c_int64(0)
    _lib.check(eng.lib.pg_prepare_features(eng._h, _stream(eng.device), Th, int(n_audio_frames),
                                           C.c_void_p(feats.data_ptr()),
                                           C.c_void_p(0 if feats0 is None else feats0.data_ptr()),
                                           C.c_void_p(0 if pitchf is None else pitchf.data_ptr()),
                                           C.c_float(protect), C.c_void_p(out.data_ptr()), C.byref(got)),
               "pg_prepare_features")
    assert got.value == p_len
    return out


def postprocess(eng: Engine, audio: torch.Tensor, src_audio: Optional[torch.Tensor] = None, src_rate: int = 16000,
                tgt_rate: int = 48000, rms_mix_rate: float = 1.0, pcm_out: Optional[torch.Tensor] = None,
                want_float: bool = False):
    """audio f32 [n] on the GPU (trimmed + concatenated) -> int16 [n] (``pcm_out``: CUDA or pinned CPU
    tensor; a fresh CUDA tensor when None), optionally also the f32 audio after change_rms."""
    audio = audio.reshape(-1).contiguous()
    n = audio.numel()
    pcm = pcm_out if pcm_out is not None else torch.empty(n, dtype=torch.int16, device=audio.device)
    fl = torch.empty(n, dtype=torch.float32, device=audio.device) if want_float else None
    if src_audio is not None:
        src_audio = src_audio.reshape(-1).to(torch.float32).contiguous()
    _lib.check(eng.lib.pg_postprocess(eng._h, _stream(eng.device), C.c_void_p(audio.data_ptr()), n,
                                      C.c_void_p(0 if src_audio is None else src_audio.data_ptr()),
                                      0 if src_audio is None else src_audio.numel(), int(src_rate), int(tgt_rate),
                                      C.c_float(rms_mix_rate), C.c_void_p(0 if fl is None else fl.data_ptr()),
                                      C.c_void_p(pcm.data_ptr())), "pg_postprocess")
    return (pcm, fl) if want_float else pcm


class ClipConverter:
    """``VC.pipeline``'s segment loop and post step as one batched device pass (pipeline.py:381-461).

    ``convert(segments, ...)`` takes the per-segment infer arguments in clip order -- what the
    reference passes to ``net_g.infer`` inside ``VC.vc`` -- decodes them as ragged batches over the
    scheduler's lanes straight into ONE clip-long device buffer (each segment's ``t_pad_tgt`` ends
    dropped on copy-out), applies change_rms / peak normalisation / int16 on the GPU and returns the
    int16 clip: the waveform crosses PCIe once, as int16 (half the fp32 bytes of the reference loop).
    """

    def __init__(self, sched: SegmentScheduler, tgt_sr: int, plan: SegmentPlan = SegmentPlan(), depth: int = 2):
        self.sched = sched
        self.tgt_sr = int(tgt_sr)
        self.t_pad_tgt = self.tgt_sr * plan.x_pad
        # `depth` clips in flight: each slot owns a clip-long f32 device buffer, a pinned int16 buffer and
        # the event that marks its post step done (the lanes wait on it before reusing the slot)
        self._slots = [{"buf": None, "pcm": None, "done": None} for _ in range(max(1, depth))]
        self._n = 0
        with torch.cuda.device(sched.device):
            self._post = torch.cuda.Stream(device=sched.device)

    def convert(self, segments: Sequence, source_audio: Optional[torch.Tensor] = None, volume_envelope: float = 1.0,
                seed: Optional[int] = None, pcm_out: Optional[torch.Tensor] = None, sync: bool = True):
        """segments: list of (phone, lengths, pitch, f0, sid) per segment, in clip order (CUDA tensors that
        are complete on the caller's current stream, or pinned CPU tensors).  source_audio: the 16 kHz clip
        (for change_rms when volume_envelope != 1).  Returns int16 [n]: ``pcm_out`` (pinned CPU or CUDA) or a
        pinned CPU tensor owned by this converter (valid until `depth` more clips were converted).
        sync=False returns without waiting (a stream of clips: the TextEncoder / flow phase of the next clip
        overlaps this clip's decoder and post step); call ``wait()`` before reading the result."""
        sched, upp, trim = self.sched, self.sched.cfg.upp, self.t_pad_tgt
        dev = torch.device("cuda", sched.device)
        frames = [int((s["phone"] if isinstance(s, dict) else s[0]).shape[-2]) for s in segments]
        counts = [T * upp - 2 * trim for T in frames]
        if min(counts) <= 0:
            raise ValueError("a segment is shorter than its two t_pad_tgt ends")
        n = sum(counts)
        slot = self._slots[self._n % len(self._slots)]
        self._n += 1
        if slot["buf"] is None or slot["buf"].numel() < n:
            slot["buf"] = torch.empty(n, dtype=torch.float32, device=dev)
        clip = slot["buf"][:n]
        views, off = [], 0
        for c in counts:
            views.append(clip[off:off + c])
            off += c
        cur = torch.cuda.current_stream(sched.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        for st in sched.streams:
            st.wait_event(ready)                       # the caller's inputs are complete
            if slot["done"] is not None:
                st.wait_event(slot["done"])            # the slot's previous clip has left its buffers
        sched.decode(segments, seed=seed, out=views, trim=trim, join=False)
        if pcm_out is None:
            if slot["pcm"] is None or slot["pcm"].numel() < n:
                slot["pcm"] = torch.empty(n, dtype=torch.int16).pin_memory()
            pcm_out = slot["pcm"][:n]
        for st in sched.streams:
            ev = torch.cuda.Event()
            ev.record(st)
            self._post.wait_event(ev)
        with torch.cuda.stream(self._post):
            src = None
            if source_audio is not None and volume_envelope != 1.0:
                src = source_audio.to(dev, torch.float32, non_blocking=True)
            postprocess(sched.engines[0], clip, src, 16000, self.tgt_sr, float(volume_envelope), pcm_out=pcm_out)
            slot["done"] = torch.cuda.Event()
            slot["done"].record(self._post)
        if sync:
            self.wait()
        return pcm_out

    def wait(self):
        """block the host until every clip handed to convert() so far is complete"""
        self._post.synchronize()
