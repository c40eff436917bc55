"""GPU parity at the sizes bench.py times (BASELINE.json configs): the launch plans there pick the
MT = 2 / 4 tiles of conv_planes_kernel / pair_planes_kernel, which the short clips of
test_gpu_parity.py never reach.  Every case compares against the CPU oracle at the BASELINE tolerance
AND asserts -- through the MT column of pg_profile_table -- that the big-tile variants really ran."""
import functools

import pytest
import torch

from conftest import snr_db
from test_gpu_parity import LATENT_REL, WAVE_MAXABS, WAVE_SNR_DB, _dev, _seg_rows, latent_err

pytestmark = pytest.mark.gpu


@functools.lru_cache(maxsize=None)
def _weights(name, seed):
    import polgen_rvc_b200 as pg
    return pg.synth_weights(pg.CONFIGS[name], seed=seed, post_std=0.02)


def _profiled_engine(cfg, sd):
    import polgen_rvc_b200 as pg
    from polgen_rvc_b200 import _lib
    return pg.Engine(cfg, pg.fold_state_dict(sd), 0, _lib.PG_FLAG_PROFILE)


def _max_mt(eng):
    """largest MT among the decoder conv launches since the last read, per channel width"""
    eng.profile_read()
    mt = {}
    for cls, cin, n, k, dil, m, cnt, ms, fl in eng.profile_table():
        if int(cls) == 0:
            mt[int(cin)] = max(mt.get(int(cin), 0), int(m))
    return mt


def _check(wave, aux, ref, tag):
    o, _, lat = ref
    assert snr_db(wave, o[:, 0]) >= WAVE_SNR_DB, tag
    assert (wave - o[:, 0]).abs().max().item() <= WAVE_MAXABS, tag
    for got, want in zip(aux, lat):
        assert latent_err(got, want) <= LATENT_REL, tag


def test_config1_size_10s_48k_vs_oracle():
    """v2-48k, T = 1000 (the 10 s row of BASELINE configs[0]/[2]), B = 1: oracle parity with MT >= 2 tiles."""
    import polgen_rvc_b200 as pg
    from oracle import rvc_oracle as orc
    cfg = pg.CONFIGS["v2-48k"]
    sd = _weights("v2-48k", 0)
    T = 1000
    inputs = pg.synth_inputs(cfg, 1, T, seed=0)
    noise = pg.synth_noise(cfg, 1, T, seed=0)
    ref = orc.infer(sd, cfg, *inputs, *noise)
    eng = _profiled_engine(cfg, sd)
    d = _dev()
    wave, aux = eng.infer(*[t.to(d) for t in inputs], noise[0].transpose(1, 2).contiguous().to(d),
                          noise[1].reshape(1, -1).contiguous().to(d), 0)
    torch.cuda.synchronize()
    _check(wave.cpu(), [aux[i].cpu().transpose(1, 2) for i in range(4)], ref, "T=1000")
    mt = _max_mt(eng)
    assert mt.get(128, 0) >= 2 and mt.get(32, 0) >= 2, mt


def test_bench_clip_segments_ragged_vs_oracle():
    """The bench workload itself (BASELINE configs[1]): the [3435, 2965]-frame segments of the 60 s clip as
    ONE ragged call -- each row against the oracle's stand-alone decode of that segment."""
    import polgen_rvc_b200 as pg
    from oracle import rvc_oracle as orc
    cfg = pg.CONFIGS["v2-48k"]
    sd = _weights("v2-48k", 0)
    frames = [3435, 2965]
    d = _dev()
    rows, raw = _seg_rows(cfg, frames, 100, d, noise=True)
    eng = _profiled_engine(cfg, sd)
    waves, auxes = eng.infer_segments(rows, seed=0, want_aux=True)
    torch.cuda.synchronize()
    mt = _max_mt(eng)
    assert mt.get(256, 0) >= 1 and mt.get(128, 0) >= 2 and mt.get(64, 0) >= 2 and mt.get(32, 0) >= 2, mt
    for (inputs, noise), w, aux, T in zip(raw, waves, auxes, frames):
        ref = orc.infer(sd, cfg, *inputs, *noise)
        _check(w.cpu()[None], [a.cpu().transpose(0, 1)[None] for a in aux], ref, f"T={T}")


def test_equal_length_batch_8x1000_vs_oracle():
    """B = 8 x T = 1000 (a sub-batch of BASELINE configs[2], 512 x 10 s): dense pg_infer against the oracle."""
    import polgen_rvc_b200 as pg
    from oracle import rvc_oracle as orc
    cfg = pg.CONFIGS["v2-48k"]
    sd = _weights("v2-48k", 0)
    B, T = 8, 1000
    inputs = pg.synth_inputs(cfg, B, T, seed=9)
    noise = pg.synth_noise(cfg, B, T, seed=9)
    ref = orc.infer(sd, cfg, *inputs, *noise)
    eng = _profiled_engine(cfg, sd)
    d = _dev()
    wave, aux = eng.infer(*[t.to(d) for t in inputs], noise[0].transpose(1, 2).contiguous().to(d),
                          noise[1].reshape(B, -1).contiguous().to(d), 0)
    torch.cuda.synchronize()
    _check(wave.cpu(), [aux[i].cpu().transpose(1, 2) for i in range(4)], ref, "B=8")
    mt = _max_mt(eng)
    assert mt.get(128, 0) >= 2 and mt.get(64, 0) >= 2, mt      # (C = 256: N = 256 fills TMEM at MT = 1)


@pytest.mark.parametrize("name", ["v2-32k", "v1-40k", "v2-40k"])
def test_other_configs_10s_vs_oracle(name):
    """BASELINE configs[0], [3], [4] at their 10 s size."""
    import polgen_rvc_b200 as pg
    from oracle import rvc_oracle as orc
    cfg = pg.CONFIGS[name]
    sd = _weights(name, 1)
    T = 1000
    inputs = pg.synth_inputs(cfg, 1, T, seed=2)
    noise = pg.synth_noise(cfg, 1, T, seed=2)
    ref = orc.infer(sd, cfg, *inputs, *noise)
    eng = _profiled_engine(cfg, sd)
    d = _dev()
    wave, aux = eng.infer(*[t.to(d) for t in inputs], noise[0].transpose(1, 2).contiguous().to(d),
                          noise[1].reshape(1, -1).contiguous().to(d), 0)
    torch.cuda.synchronize()
    _check(wave.cpu(), [aux[i].cpu().transpose(1, 2) for i in range(4)], ref, name)
    assert max(_max_mt(eng).values()) >= 2


def test_padded_launch_shape_equals_unpadded_twin():
    """T is padded up to a bucket (1000 -> 1024, 257 -> 288) with the true T as a hard end read from device
    memory: the result must equal the PG_FLAG_NO_PAD twin that launches at exactly T (up to the attention's
    key-split count, which depends on the launched T: fp32 reassociation only)."""
    import polgen_rvc_b200 as pg
    from polgen_rvc_b200 import _lib
    cfg = pg.CONFIGS["v2-48k"]
    sd = _weights("v2-48k", 0)
    folded = pg.fold_state_dict(sd)
    d = _dev()
    a_eng, b_eng = pg.Engine(cfg, folded, 0), pg.Engine(cfg, folded, 0, _lib.PG_FLAG_NO_PAD)
    for B, T in ((1, 257), (2, 1000), (1, 33)):
        assert a_eng.padded_frames(T) > T and b_eng.padded_frames(T) == T
        inp = [t.to(d) for t in pg.synth_inputs(cfg, B, T, seed=T)]
        wa, xa = a_eng.infer(*inp, None, None, 5)
        wb, xb = b_eng.infer(*inp, None, None, 5)
        torch.cuda.synchronize()
        assert wa.shape == wb.shape == (B, T * cfg.upp)
        assert snr_db(wa.cpu(), wb.cpu()) >= 80.0, (B, T)
        for i in range(4):
            assert latent_err(xa[i].cpu(), xb[i].cpu()) <= 1e-4


def test_graph_buckets_replay_distinct_lengths_and_lru():
    """A stream of segments with DISTINCT lengths (what the reference pipeline produces) replays one CUDA
    graph per bucket; the cache is bounded (LRU).  Through the drop-in call on torch's default stream."""
    import polgen_rvc_b200 as pg
    from polgen_rvc_b200 import _lib
    cfg = pg.CONFIGS["v2-48k"]
    sd = _weights("v2-48k", 0)
    d = _dev()
    net = pg.Synthesizer(*cfg.ctor_args(), use_f0=1, input_dim=cfg.input_dim, is_half=False)
    net.load_state_dict(sd, strict=False)
    net.eval().to(d)
    twin = pg.Engine(cfg, pg.fold_state_dict(sd), 0, _lib.PG_FLAG_NO_GRAPHS)
    lengths = [100, 111, 97, 120, 128, 105]            # all in the (96, 128] bucket
    assert len({net.engine().padded_frames(T) for T in lengths}) == 1
    for n, T in enumerate(lengths):
        phone, ln, pitch, f0, sid = [t.to(d) for t in pg.synth_inputs(cfg, 1, T, seed=n)]
        torch.manual_seed(n)
        audio = net.infer(phone, ln, pitch, f0, sid)[0][0, 0].data.cpu().float()   # pipeline.py:275-279
        torch.manual_seed(n)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            want, _ = twin.infer(phone, ln, pitch, f0, sid, None, None, seed, want_aux=False)
        st.synchronize()
        assert audio.shape == (T * cfg.upp,)
        assert torch.equal(audio, want.cpu()[0]), T     # replayed graph == directly launched twin
    eng = net.engine()
    assert eng.graph_count() == 1                        # one graph served all six lengths
    eng.set_graph_cache(2)
    net.infer(*[t.to(d) for t in pg.synth_inputs(cfg, 1, 400, seed=400)])   # grow the workspace first: a growth drops all graphs
    for T in (400, 40, 200, 300):                        # four more buckets, twice each -> captured, then trimmed
        inp = [t.to(d) for t in pg.synth_inputs(cfg, 1, T, seed=T)]
        for _ in range(2):
            net.infer(*inp)
    torch.cuda.synchronize()
    assert eng.graph_count() == 2


def test_two_devices_in_one_process():
    """cudaFuncSetAttribute / SM count are per device: an engine on cuda:1 after one on cuda:0 must work."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import polgen_rvc_b200 as pg
    cfg = pg.CONFIGS["v2-48k"]
    sd = _weights("v2-48k", 0)
    folded = pg.fold_state_dict(sd)
    outs = []
    for dev in (0, 1):
        eng = pg.Engine(cfg, folded, dev)
        with torch.cuda.device(dev):
            inp = [t.to(f"cuda:{dev}") for t in pg.synth_inputs(cfg, 1, 300, seed=1)]
            w, _ = eng.infer(*inp, None, None, 3, want_aux=False)
            torch.cuda.synchronize(dev)
            outs.append(w.cpu())
    assert torch.equal(outs[0], outs[1])


def test_time_tiled_decoder_matches_whole_segment():
    """GeneratorNSF decoded as time tiles with a 16-frame halo (what each of N GPUs
    would do for one long segment) equals the whole-segment decode, and the oracle."""
    import polgen_rvc_b200 as pg
    from oracle import rvc_oracle as orc
    cfg = pg.CONFIGS["v2-48k"]
    sd = _weights("v2-48k", 0)
    eng = pg.Engine(cfg, pg.fold_state_dict(sd), 0)
    d = _dev()
    T = 700
    inputs = pg.synth_inputs(cfg, 1, T, seed=3)
    noise = pg.synth_noise(cfg, 1, T, seed=3)
    dev_in = [t.to(d) for t in inputs]
    ez, es = noise[0].transpose(1, 2).contiguous().to(d), noise[1].reshape(1, -1).contiguous().to(d)
    whole, _ = eng.infer(*dev_in, ez, es, 0, want_aux=False)
    for n_tiles in (2, 5, 8):
        tiled = pg.TimeTiledDecoder(eng).decode(*dev_in, n_tiles=n_tiles, eps_zp=ez, eps_src=es)
        torch.cuda.synchronize()
        assert tiled.shape == (T * cfg.upp,)
        assert snr_db(tiled.cpu(), whole.cpu()[0]) >= 90.0, n_tiles
    o = orc.infer(sd, cfg, *inputs, *noise)[0]
    assert snr_db(tiled.cpu(), o[0, 0]) >= WAVE_SNR_DB
