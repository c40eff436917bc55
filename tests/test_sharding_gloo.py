"""world_size-2 gloo run of the segment-sharded decode driver on CPU: every rank decodes its
share with a stand-in decode function and rank 0 receives all waveforms in order.  (The decode
itself has no collective -- SURVEY.md 8(e); this covers the N>1 host logic.)"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from polgen_rvc_b200 import segments as seg
    lengths = [400, 240, 100, 100, 370, 55, 240]
    calls = []

    def decode(indices):
        calls.append(list(indices))
        return {i: np.full(lengths[i] * 3, float(i), dtype=np.float32) + rank * 0.0 for i in indices}

    res = seg.decode_sharded(decode, lengths, rank, world)
    mine = seg.plan_shards(lengths, world)[rank]
    assert calls == [mine]
    t = torch.tensor([float(sum(lengths[i] for i in mine))])
    dist.all_reduce(t)
    assert int(t.item()) == sum(lengths)
    # equal-length batch (BASELINE configs[2]): tensor gather of unequal shards, rank 0 gets segment order
    n_seg, n_samples = 7, 5
    bins = seg.plan_shards([1000] * n_seg, world)
    local = torch.stack([torch.full((n_samples,), float(i)) for i in bins[rank]])
    full = seg.gather_waveforms(local, bins, n_samples, rank, world)
    if rank == 0:
        assert res is not None and len(res) == len(lengths)
        for i, w in enumerate(res):
            assert w.shape[0] == lengths[i] * 3 and float(w[0]) == float(i)
        assert full.shape == (n_seg, n_samples)
        assert torch.equal(full[:, 0], torch.arange(n_seg, dtype=torch.float32))
        open(out_path, "w").write("ok")
    else:
        assert res is None and full is None
    dist.barrier()
    dist.destroy_process_group()


def test_decode_sharded_two_ranks(tmp_path):
    out = str(tmp_path / "ok.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


class _FakeEngine:
    """CPU stand-in with the entry points TimeTiledDecoder uses: the 'decoder' is a causal-free local filter
    (7-frame receptive field) so tiling with a halo must reproduce the untiled result exactly."""

    class cfg:
        upp = 4
        inter_channels = 6

    def text_encoder(self, phone, lengths, pitch):
        m = phone[..., :6].clone()
        return m, torch.zeros_like(m)

    def flow_reverse(self, z_p, lengths, sid):
        return z_p * 2.0

    def source(self, f0, eps_src, seed):
        return f0.repeat_interleave(4, dim=1), None

    def generator(self, z, src, sid):
        x = z.sum(-1)                                                  # [B][T]
        y = torch.nn.functional.conv1d(x[:, None], torch.ones(1, 1, 7) / 7, padding=3)[:, 0]
        return y.repeat_interleave(4, dim=1) + 0.1 * src


def _tile_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import polgen_rvc_b200 as pg
    T = 211
    g = torch.Generator().manual_seed(3)
    phone = torch.randn(1, T, 8, generator=g)
    f0 = torch.rand(1, T, generator=g)
    eps = torch.zeros(1, T, 6)
    args = (phone, torch.tensor([T]), torch.zeros(1, T, dtype=torch.long), f0, torch.tensor([0]))
    whole = pg.TimeTiledDecoder(_FakeEngine(), 0, 1).decode(*args, n_tiles=1, eps_zp=eps)
    tiled = pg.TimeTiledDecoder(_FakeEngine(), rank, world).decode(*args, n_tiles=5, eps_zp=eps)
    if rank == 0:
        assert tiled.shape == whole.shape == (T * 4,)
        assert torch.allclose(tiled, whole, atol=1e-6)
        open(out_path, "w").write("ok")
    else:
        assert tiled is None
    dist.barrier()
    dist.destroy_process_group()


def test_time_tiled_decoder_two_ranks(tmp_path):
    """SURVEY 8(f) rank 4 host logic over gloo: tiles rank::world, halo windows, gather on rank 0."""
    out = str(tmp_path / "ok.txt")
    mp.spawn(_tile_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"
