"""GPU diagnostics (developer tool, run under gpurun): per-tap parity of the CUDA
path against the oracle, op-level tcgen05-vs-CUDA-core conv checks with timing,
and quick end-to-end timings.  Each stage runs in its own subprocess so a
trapping kernel cannot take the other results with it.

    python tools/gpu_diag.py all            # everything, JSON lines to gpurun_out/diag.jsonl
    python tools/gpu_diag.py taps simt|umma [config] [T] [B]
    python tools/gpu_diag.py convop
    python tools/gpu_diag.py time [config] [T] [B]
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def emit(rec):
    os.makedirs(OUT, exist_ok=True)
    line = json.dumps(rec)
    print(line, flush=True)
    with open(os.path.join(OUT, "diag.jsonl"), "a") as f:
        f.write(line + "\n")


def snr_db(a, b):
    import torch
    a = a.double().flatten()
    b = b.double().flatten()
    return float(10 * torch.log10((b ** 2).sum() / ((a - b) ** 2).sum().clamp_min(1e-300)))


def taps(mode, name="v2-40k", T=24, B=1, seed=0):  # noqa: C901
    import torch
    import polgen_rvc_b200 as pg
    from polgen_rvc_b200 import _lib
    from oracle import rvc_oracle as orc
    cfg = pg.CONFIGS[name]
    sd = pg.synth_weights(cfg, seed=seed)
    phone, lengths, pitch, f0, sid = pg.synth_inputs(cfg, B, T, seed=seed)
    eps_zp, eps_src = pg.synth_noise(cfg, B, T, seed=seed)
    ot = {}
    o, m, (z, z_p, m_p, logs_p) = orc.infer(sd, cfg, phone, lengths, pitch, f0, sid, eps_zp, eps_src, taps=ot)
    flags = _lib.PG_FLAG_KEEP_TAPS | (_lib.PG_FLAG_FORCE_SIMT if mode == "simt" else 0)
    eng = pg.Engine(cfg, pg.fold_state_dict(sd), 0, flags)
    d = torch.device("cuda:0")
    wave, aux = eng.infer(phone.to(d), lengths.to(d), pitch.to(d), f0.to(d), sid.to(d),
                          eps_zp.transpose(1, 2).contiguous().to(d), eps_src.reshape(B, -1).contiguous().to(d), 0)
    torch.cuda.synchronize()
    rec = {"stage": "taps", "mode": mode, "config": name, "T": T, "B": B, "launches": eng.launch_count()}
    res = {}

    def cmp(key, got_btc, want_bct):
        want = want_bct.transpose(1, 2)
        got = got_btc.cpu()
        res[key] = {"maxabs": float((got - want).abs().max()), "ref_absmax": float(want.abs().max()),
                    "snr_db": round(snr_db(got, want), 2)}

    for i, (k, w) in enumerate([("z", z), ("z_p", z_p), ("m_p", m_p), ("logs_p", logs_p)]):
        cmp(k, aux[i], w)
    names = ["enc.x0"] + [f"enc.layer{i}" for i in range(cfg.n_layers)] + \
            [f"flow.{f}" for f in range(cfg.flow_n_flows)] + ["dec.conv_pre"]
    for i in range(len(cfg.upsample_rates)):
        names += [f"dec.ups{i}", f"dec.stage{i}"]
    for k in names:
        try:
            cmp(k, eng.fetch_tap(k), ot[k])
        except Exception as e:  # noqa: BLE001
            res[k] = {"error": str(e)}
    src = eng.fetch_tap("source").cpu().reshape(B, -1)
    res["source"] = {"maxabs": float((src - ot["source"].reshape(B, -1)).abs().max())}
    res["wave"] = {"maxabs": float((wave.cpu() - o[:, 0]).abs().max()), "ref_absmax": float(o.abs().max()),
                   "snr_db": round(snr_db(wave.cpu(), o[:, 0]), 2)}
    rec["taps"] = res
    emit(rec)


def snrscan():
    """wave SNR over configs / seeds / lengths, tcgen05 and CUDA-core paths."""
    import torch
    import polgen_rvc_b200 as pg
    from polgen_rvc_b200 import _lib
    from oracle import rvc_oracle as orc
    d = torch.device("cuda:0")
    for name, T, seed in [("v1-40k", 30, 8), ("v1-40k", 30, 9), ("v2-40k", 30, 8), ("v2-48k", 30, 8),
                          ("v2-32k", 30, 8), ("v1-40k", 13, 3), ("v2-40k", 100, 4), ("v2-48k", 100, 5),
                          ("v1-40k", 100, 6), ("v2-32k", 100, 7)]:
        cfg = pg.CONFIGS[name]
        sd = pg.synth_weights(cfg, seed=seed)
        inp = pg.synth_inputs(cfg, 1, T, seed=seed)
        eps_zp, eps_src = pg.synth_noise(cfg, 1, T, seed=seed)
        o = orc.infer(sd, cfg, *inp, eps_zp, eps_src)[0][:, 0]
        rec = {"stage": "snrscan", "config": name, "T": T, "seed": seed, "o_rms": float(o.pow(2).mean().sqrt())}
        for mode, flags in (("umma", 0), ("simt", _lib.PG_FLAG_FORCE_SIMT)):
            eng = pg.Engine(cfg, pg.fold_state_dict(sd), 0, flags)
            wave, _ = eng.infer(*[t.to(d) for t in inp], eps_zp.transpose(1, 2).contiguous().to(d),
                                eps_src.reshape(1, -1).contiguous().to(d), 0)
            rec[mode] = {"snr_db": round(snr_db(wave.cpu(), o), 2), "maxabs": float((wave.cpu() - o).abs().max())}
            eng.close()
        emit(rec)


def convop():
    import torch
    from polgen_rvc_b200 import _lib
    lib = _lib.load()
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    for (Cc, K, dil, L, B) in [(128, 7, 3, 1000, 2), (64, 11, 5, 2048, 1), (32, 11, 3, 4096, 2),
                               (256, 3, 1, 600, 1),
                               (128, 3, 1, 120000, 1), (128, 7, 1, 120000, 1), (128, 11, 5, 120000, 1),
                               (64, 3, 1, 240000, 1), (64, 7, 1, 240000, 1), (64, 11, 5, 240000, 1),
                               (32, 3, 1, 480000, 1), (32, 7, 1, 480000, 1), (32, 11, 5, 480000, 1),
                               (256, 3, 1, 12000, 1), (256, 7, 1, 12000, 1), (256, 11, 5, 12000, 1)]:
        x = (torch.randn(B, L, Cc, generator=g) * 1.0).half()
        w = torch.randn(Cc, Cc, K, generator=g) * (1.0 / (Cc * K) ** 0.5)
        bias = torch.randn(Cc, generator=g) * 0.1
        res = (torch.randn(B, L, Cc, generator=g)).half()
        xd, rd = x.to(d), res.to(d)
        ref = torch.nn.functional.conv1d(
            torch.nn.functional.leaky_relu(xd.float(), 0.1).transpose(1, 2), w.half().float().to(d), bias.to(d),
            dilation=dil, padding=(K * dil - dil) // 2).transpose(1, 2) + rd.float()
        rec = {"stage": "convop", "C": Cc, "K": K, "dil": dil, "L": L, "B": B, "mt": os.environ.get("PG_UMMA_MT", "auto")}
        for impl, nm in [(0, "simt"), (1, "umma"), (2, "planes")]:
            y = torch.zeros(B, L, Cc, device=d, dtype=torch.half)
            ms = C.c_float(0)
            iters = 5
            rc = lib.pg_op_conv1d_f16(0, impl, B, L, Cc, Cc, K, dil, C.c_void_p(xd.data_ptr()),
                                      C.c_void_p(w.contiguous().data_ptr()), C.c_void_p(bias.data_ptr()),
                                      C.c_float(0.1), C.c_float(1.0), C.c_void_p(rd.data_ptr()),
                                      C.c_void_p(y.data_ptr()), iters, C.byref(ms))
            if rc != 0:
                rec[nm] = {"error": lib.pg_last_error().decode()}
                continue
            torch.cuda.synchronize()
            err = (y.float() - ref).abs().max().item()
            fl = 2.0 * B * L * Cc * Cc * K
            rec[nm] = {"maxabs": err, "ref_absmax": ref.abs().max().item(), "ms": ms.value / iters,
                       "tflops": fl / (ms.value / iters * 1e-3) / 1e12}
        emit(rec)


def timeit(name="v2-48k", T=1000, B=1, mode="umma"):
    import torch
    import polgen_rvc_b200 as pg
    from polgen_rvc_b200 import _lib
    cfg = pg.CONFIGS[name]
    sd = pg.synth_weights(cfg, seed=0)
    phone, lengths, pitch, f0, sid = pg.synth_inputs(cfg, B, T, seed=0)
    eng = pg.Engine(cfg, pg.fold_state_dict(sd), 0, _lib.PG_FLAG_FORCE_SIMT if mode == "simt" else 0)
    d = torch.device("cuda:0")
    args = [t.to(d) for t in (phone, lengths, pitch, f0, sid)]
    for _ in range(2):
        eng.infer(*args, None, None, 0, want_aux=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        eng.infer(*args, None, None, 0, want_aux=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    emit({"stage": "time", "mode": mode, "config": name, "T": T, "B": B, "ms": ms,
          "audio_s_per_s": B * T / 100.0 / (ms * 1e-3), "launches": eng.launch_count(),
          "ws_MB": eng.workspace_bytes(B, T) / 2 ** 20})


def run_jobs(jobs):
    for j in jobs:
        env = dict(os.environ)
        if isinstance(j, tuple):
            env.update(j[1])
            j = j[0]
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + j, timeout=420,
                               capture_output=True, text=True, env=env)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                emit({"stage": "job_failed", "job": j, "rc": r.returncode, "stderr": r.stderr[-3000:],
                      "stdout_tail": r.stdout[-1500:]})
        except subprocess.TimeoutExpired:
            emit({"stage": "job_timeout", "job": j})
        print(f"# job {j} took {time.time() - t0:.1f}s", flush=True)


def run_tune():
    run_jobs([
        ["taps", "umma", "v2-48k", "300", "1"],
        ["convop"],
        (["convop"], {"PG_UMMA_MT": "2"}),
        ["time", "v2-48k", "1000", "1", "umma"],
        ["time", "v2-48k", "4000", "1", "umma"],
    ])


def run_all():
    jobs = [
        ["taps", "simt", "v2-40k", "24", "1"],
        ["taps", "simt", "v2-32k", "12", "2"],
        ["convop"],
        ["taps", "umma", "v2-40k", "24", "1"],
        ["taps", "umma", "v2-48k", "300", "1"],
        ["taps", "simt", "v2-48k", "300", "1"],
        ["time", "v2-48k", "1000", "1", "simt"],
        ["time", "v2-48k", "1000", "1", "umma"],
        ["time", "v2-48k", "4000", "1", "umma"],
    ]
    run_jobs(jobs)


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "all"
    a = sys.argv[2:]
    if cmd == "all":
        run_all()
    elif cmd == "tune":
        run_tune()
    elif cmd == "taps":
        taps(a[0], a[1] if len(a) > 1 else "v2-40k", int(a[2]) if len(a) > 2 else 24,
             int(a[3]) if len(a) > 3 else 1, int(a[4]) if len(a) > 4 else 0)
    elif cmd == "snrscan":
        snrscan()
    elif cmd == "convop":
        convop()
    elif cmd == "time":
        timeit(a[0] if a else "v2-48k", int(a[1]) if len(a) > 1 else 1000, int(a[2]) if len(a) > 2 else 1,
               a[3] if len(a) > 3 else "umma")
