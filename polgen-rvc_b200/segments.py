"""Segment sharding: the data-parallel unit of the decode.

The reference converts a clip as a sequential loop of independent ``infer``
calls, one per silence-split segment (``rvc/infer/pipeline.py:329-348`` picks
the cut points, ``:381-447`` runs the loop, ``:397`` trims ``t_pad_tgt`` from
both ends, ``:449`` concatenates).  Nothing flows between segments, so they
shard across GPUs with no collective inside the decode:
segments are bin-packed over ranks by length, every rank decodes its share
with its own engine (weights replicated), and the host gathers the trimmed
waveforms in order.

This module restates only the segmentation arithmetic and the scheduling; the
HuBERT / F0 / faiss stages that produce the per-segment features stay in the
reference pipeline.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

SAMPLE_RATE = 16000     # pipeline.py:75
WINDOW = 160            # pipeline.py:76  (100 frames per second)


@dataclass(frozen=True)
class SegmentPlan:
    """x_pad/x_query/x_center/x_max seconds -> sample counts (pipeline.py:65-84)."""
    x_pad: int = 1
    x_query: int = 6
    x_center: int = 38
    x_max: int = 41        # (1, 6, 38, 41) on CUDA devices, infer.py:41-45

    @property
    def t_pad(self): return SAMPLE_RATE * self.x_pad
    @property
    def t_pad2(self): return 2 * self.t_pad
    @property
    def t_query(self): return SAMPLE_RATE * self.x_query
    @property
    def t_center(self): return SAMPLE_RATE * self.x_center
    @property
    def t_max(self): return SAMPLE_RATE * self.x_max


def split_points(audio: np.ndarray, plan: SegmentPlan = SegmentPlan()) -> List[int]:
    """Cut positions ``opt_ts`` (16 kHz samples) exactly as pipeline.py:330-344: if the clip
    (+window) exceeds t_max, every t_center samples take the minimum of the 160-tap moving
    |sum| envelope within +-t_query."""
    audio = np.asarray(audio, dtype=np.float64)
    pad = np.pad(audio, (WINDOW // 2, WINDOW // 2), mode="reflect")
    cuts: List[int] = []
    if pad.shape[0] > plan.t_max:
        csum = np.concatenate([[0.0], np.cumsum(pad)])
        env = np.abs(csum[WINDOW:WINDOW + audio.shape[0]] - csum[:audio.shape[0]])
        # (the reference accumulates audio_pad[i : i - window] for i in range(window))
        for t in range(plan.t_center, audio.shape[0], plan.t_center):
            lo, hi = t - plan.t_query, t + plan.t_query
            seg = env[lo:hi]
            cuts.append(lo + int(np.where(seg == seg.min())[0][0]))
    return cuts


def segment_frames(n_samples: int, cuts: Sequence[int], plan: SegmentPlan = SegmentPlan()) -> List[Tuple[int, int]]:
    """(start_frame, n_frames) of every segment inside the t_pad-reflect-padded clip
    (pipeline.py:381-447): segment i spans audio_pad[s : t + t_pad2 + window], its pitch
    slice is [s // window : (t + t_pad2) // window], and the last one runs to the end."""
    p_len = (n_samples + 2 * plan.t_pad) // WINDOW
    out: List[Tuple[int, int]] = []
    s = 0
    for t in cuts:
        t = t // WINDOW * WINDOW
        out.append((s // WINDOW, (t + plan.t_pad2) // WINDOW - s // WINDOW))
        s = t
    out.append((s // WINDOW, p_len - s // WINDOW))
    return out


def trim_and_concat(waves: Sequence[np.ndarray], tgt_sr: int, plan: SegmentPlan = SegmentPlan()) -> np.ndarray:
    """pipeline.py:397,447,449: drop t_pad_tgt samples at both ends of every segment, concatenate."""
    t_pad_tgt = tgt_sr * plan.x_pad
    return np.concatenate([np.asarray(w)[t_pad_tgt:len(w) - t_pad_tgt] for w in waves])


def plan_shards(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-first greedy bin packing of segment indices over ``world`` ranks (cost ~ frames).
    Deterministic: ties go to the lowest rank; each rank's list is returned in index order."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    bins: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(b) for b in bins]


def batch_equal_lengths(indices: Sequence[int], lengths: Sequence[int], max_batch: int) -> List[List[int]]:
    """Group a rank's segments into equal-length batches (the generator ignores x_mask after its
    input, so padded ragged batches would not reproduce the B=1 results -- SURVEY.md H6)."""
    by_len: Dict[int, List[int]] = {}
    for i in indices:
        by_len.setdefault(int(lengths[i]), []).append(i)
    out: List[List[int]] = []
    for _, idx in sorted(by_len.items()):
        for k in range(0, len(idx), max_batch):
            out.append(idx[k:k + max_batch])
    return out


def decode_sharded(decode_fn: Callable[[List[int]], Dict[int, np.ndarray]], lengths: Sequence[int],
                   rank: int = 0, world: int = 1, group=None) -> Optional[List[np.ndarray]]:
    """Run ``decode_fn`` on this rank's share and gather all waveforms on rank 0, in segment order.

    ``decode_fn(indices) -> {index: waveform}`` decodes locally (one engine per GPU).  The only
    communication is the host-side gather of the results (``torch.distributed.gather_object``);
    the decode itself needs no collective.  Returns the ordered list on rank 0, None elsewhere."""
    mine = plan_shards(lengths, world)[rank]
    local = decode_fn(mine)
    if set(local) != set(mine):
        raise RuntimeError("decode_fn must return exactly the requested segments")
    if world == 1:
        return [local[i] for i in range(len(lengths))]
    import torch.distributed as dist
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0, group=group)
    if rank != 0:
        return None
    merged: Dict[int, np.ndarray] = {}
    for part in gathered:
        merged.update(part)
    if len(merged) != len(lengths):
        raise RuntimeError("missing segments after gather")
    return [merged[i] for i in range(len(lengths))]


def gather_waveforms(local, owner_of: Sequence[int], n_samples: int, rank: int = 0, world: int = 1, group=None):
    """Device-side gather of equal-length waveforms for a sharded batch (the one
    communication step of the path, after the decode): ``local`` is this rank's [n_local][n_samples]
    CUDA tensor holding its segments in ``plan_shards`` order; rank 0 receives a [n_total][n_samples]
    tensor in SEGMENT order (one ``torch.distributed.gather`` of equal-size, zero-padded shards over
    NCCL / NVLink), other ranks get None.  ``owner_of``: the plan_shards bins."""
    import torch
    n_total = sum(len(b) for b in owner_of)
    if world == 1:
        out = torch.empty(n_total, n_samples, dtype=local.dtype, device=local.device)
        out[torch.as_tensor(owner_of[0], device=local.device)] = local
        return out
    import torch.distributed as dist
    width = max(len(b) for b in owner_of)
    send = torch.zeros(width, n_samples, dtype=local.dtype, device=local.device)
    send[:local.shape[0]] = local
    recv = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
    dist.gather(send, recv, dst=0, group=group)
    if rank != 0:
        return None
    out = torch.empty(n_total, n_samples, dtype=local.dtype, device=local.device)
    for r, idx in enumerate(owner_of):
        if idx:
            out[torch.as_tensor(idx, device=local.device)] = recv[r][:len(idx)]
    return out
