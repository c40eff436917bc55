"""Intra-segment time tiling of GeneratorNSF across GPUs.

A clip with fewer silence-split segments than GPUs (BASELINE configs[1]: 2 segments) cannot use more
than that many GPUs through segment sharding alone.  The decoder (``GeneratorNSF.forward``,
``rvc/lib/algorithm/nsf.py:120-144``) is purely convolutional, so a segment's frames can be cut
into tiles that are decoded independently, each with a halo of context frames on both sides that
covers the decoder's receptive field (conv_pre 3 frames + the first stage's ResBlocks
(k-1)/2*(1+3+5) + 3*(k-1)/2 = 60 samples at 10-12 samples per frame + the later stages' fractions
of a frame: < 11 frames for every BASELINE config; HALO = 16).  The TextEncoder (global attention)
and the flow run whole on every rank -- 2.5 % of the FLOPs, deterministic, so no broadcast is needed
-- as does the harmonic source, whose phase is a prefix sum from the start of the segment; each rank
then decodes its tiles with ``pg_generator`` on the window [a - HALO, b + HALO) and keeps [a, b).
At the true ends of the segment the window is clipped, so the zero padding there is the real one.
The only communication is the gather of the finished waveform tiles.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

HALO = 16   # frames of context on each side of a tile (>= the decoder's receptive field, see above)


def plan_tiles(T: int, n_tiles: int, align: int = 8) -> List[Tuple[int, int]]:
    """[a, b) frame ranges of ``n_tiles`` near-equal tiles covering [0, T); interior cuts on multiples of
    ``align`` frames.  Fewer tiles come back when T is too short to give every tile `align` frames."""
    n_tiles = max(1, min(int(n_tiles), max(1, T // max(1, align))))
    cuts = [0]
    for i in range(1, n_tiles):
        c = (T * i // n_tiles) // align * align
        if c > cuts[-1]:
            cuts.append(c)
    cuts.append(T)
    return [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]


class TimeTiledDecoder:
    """One segment decoded as time tiles, this rank taking tiles ``rank::world``.

    ``decode(...)`` returns, on rank 0, the whole waveform [T*upp] (fp32, on the GPU); other ranks get
    None.  With ``world == 1`` the tiles run one after another on this GPU (same code path, used by the
    parity test); results match the untiled ``Engine.infer`` to rounding."""

    def __init__(self, engine, rank: int = 0, world: int = 1, group=None, halo: int = HALO):
        self.eng, self.rank, self.world, self.group, self.halo = engine, int(rank), int(world), group, int(halo)

    def latents(self, phone, lengths, pitch, f0, sid, eps_zp=None, eps_src=None, seed: int = 0):
        """TextEncoder -> reparameterisation -> flow -> source, whole segment (synthesizers.py:172-183)."""
        eng = self.eng
        m_p, logs_p = eng.text_encoder(phone, lengths, pitch)
        T = phone.shape[1]
        mask = (torch.arange(T, device=phone.device)[None, :] < lengths[:, None]).to(torch.float32)[:, :, None]
        if eps_zp is None:
            g = torch.Generator(device=phone.device).manual_seed(int(seed) & (2 ** 62 - 1))
            eps_zp = torch.randn(m_p.shape, generator=g, device=phone.device, dtype=torch.float32)
        z_p = (m_p + torch.exp(logs_p) * eps_zp * 0.66666) * mask
        z = eng.flow_reverse(z_p.contiguous(), lengths, sid) * mask
        src, _ = eng.source(f0, eps_src, seed)
        return z.contiguous(), src

    def decode(self, phone, lengths, pitch, f0, sid, n_tiles: Optional[int] = None, eps_zp=None, eps_src=None,
               seed: int = 0):
        if phone.shape[0] != 1:
            raise ValueError("time tiling decodes one segment (B = 1) at a time")
        eng, upp, h = self.eng, self.eng.cfg.upp, self.halo
        T = phone.shape[1]
        tiles = plan_tiles(T, n_tiles if n_tiles is not None else self.world)
        z, src = self.latents(phone, lengths, pitch, f0, sid, eps_zp, eps_src, seed)
        mine = tiles[self.rank::self.world]
        width = max(b - a for a, b in tiles) * upp
        out = torch.zeros(max(1, len(tiles[0::self.world])), width, device=phone.device, dtype=torch.float32)
        for k, (a, b) in enumerate(mine):
            lo, hi = max(0, a - h), min(T, b + h)
            w = eng.generator(z[:, lo:hi].contiguous(), src[:, lo * upp:hi * upp].contiguous(), sid)
            out[k, :(b - a) * upp] = w[0, (a - lo) * upp:(b - lo) * upp]
        return self._gather(out, tiles, T * upp)

    def _gather(self, out, tiles: Sequence[Tuple[int, int]], n_samples: int):
        upp = self.eng.cfg.upp
        if self.world == 1:
            parts = [out]
        else:
            import torch.distributed as dist
            parts = [torch.empty_like(out) for _ in range(self.world)] if self.rank == 0 else None
            dist.gather(out, parts, dst=0, group=self.group)
            if self.rank != 0:
                return None
        wave = torch.empty(n_samples, device=out.device, dtype=torch.float32)
        for i, (a, b) in enumerate(tiles):
            wave[a * upp:b * upp] = parts[i % self.world][i // self.world, :(b - a) * upp]
        return wave
