"""Drop-in host class for the reference ``Synthesizer``.

Mirrors ``rvc/lib/algorithm/synthesizers.py:13-188`` at the module-object
boundary the callers use:

* ``Synthesizer(*cpt["config"], use_f0=..., input_dim=..., is_half=...)``
  (``rvc/infer/infer.py:92-97``)
* ``del net_g.enc_q``; ``load_state_dict(cpt["weight"], strict=False)``;
  ``.eval().to(device)``; ``.half()`` / ``.float()`` (``infer.py:99-102``)
* ``net_g.infer(feats, p_len, pitch, pitchf, sid)`` ->
  ``(o, x_mask, (z, z_p, m_p, logs_p))`` (``rvc/infer/pipeline.py:275``)

The torch modules below only HOLD parameters under the reference's state-dict
names (legacy ``weight_g/weight_v`` checkpoints load through torch's
weight-norm compat hook); none of their ``forward`` methods is ever called.
All compute goes through the C ABI (``include/polgen_rvc.h``) into the
hand-written sm_100a kernels.  There is no CPU path: ``infer`` on a non-CUDA
tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
from torch import nn
from torch.nn.utils.parametrizations import weight_norm

from . import _lib
from .configs import SynthConfig, config_from_ctor


# --------------------------------------------------------------------------
# weight-norm fold (done once at load; the reference recomputes it every
# forward because remove_weight_norm is never called -- SURVEY.md K12)
# --------------------------------------------------------------------------
def fold_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Return plain fp32 ``<module>.weight`` tensors: w = g * v / ||v|| with the
    norm over every dim but 0 (Cout for Conv1d, Cin for ConvTranspose1d)."""
    g_sfx = (".weight_g", ".parametrizations.weight.original0")
    v_sfx = (".weight_v", ".parametrizations.weight.original1")
    out: Dict[str, torch.Tensor] = {}
    for key, val in sd.items():
        if key.endswith(g_sfx):
            continue
        hit = next((i for i, s in enumerate(v_sfx) if key.endswith(s)), None)
        if hit is None:
            out[key] = val.detach().to("cpu", torch.float32).contiguous()
            continue
        base = key[: -len(v_sfx[hit])]
        v = val.detach().to("cpu", torch.float32)
        g = sd[base + g_sfx[hit]].detach().to("cpu", torch.float32)
        flat = v.reshape(v.shape[0], -1)
        scale = g.reshape(-1) / flat.norm(dim=1)
        out[base + ".weight"] = (flat * scale[:, None]).reshape(v.shape).contiguous()
    return out


# --------------------------------------------------------------------------
# Engine: one C-ABI handle on one GPU
# --------------------------------------------------------------------------
class Engine:
    """Owns a ``pg_handle``.  ``weights`` is a folded fp32 CPU state dict."""

    def __init__(self, cfg: SynthConfig, weights: Dict[str, torch.Tensor], device: int = 0,
                 flags: int = 0):
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = int(device)
        self._h = C.c_void_p()
        pc = _lib.make_config(cfg, flags)
        _lib.check(self.lib.pg_create(C.byref(pc), self.device, C.byref(self._h)), "pg_create")
        try:
            for name, t in weights.items():
                if name.startswith("enc_q."):
                    continue
                t = t.detach().to("cpu", torch.float32).contiguous()
                shape = (C.c_int64 * max(1, t.dim()))(*t.shape)
                _lib.check(self.lib.pg_load_tensor(self._h, name.encode(), C.c_void_p(t.data_ptr()),
                                                   shape, t.dim(), _lib.PG_F32), f"pg_load_tensor({name})")
            _lib.check(self.lib.pg_finalize(self._h), "pg_finalize")
        except Exception:
            self.close()
            raise

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.pg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]):
        return C.c_void_p(0 if t is None else t.data_ptr())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self):
        return torch.device("cuda", self.device)

    def workspace_bytes(self, B: int, T: int) -> int:
        return int(self.lib.pg_workspace_bytes(self._h, B, T))

    def launch_count(self) -> int:
        return int(self.lib.pg_launch_count(self._h))

    def graph_count(self) -> int:
        """(B, T) shapes currently held as instantiated CUDA graphs (replayed from their 2nd call on)"""
        return int(self.lib.pg_graph_count(self._h))

    def profile_read(self):
        """PG_FLAG_PROFILE: {class: (ms, flops, launches)} of the conv launches since the last read."""
        ms, fl, n = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_int64 * 3)()
        _lib.check(self.lib.pg_profile_read(self._h, ms, fl, n), "pg_profile_read")
        return {"conv_planes": (ms[0], fl[0], int(n[0])), "conv_simt": (ms[1], fl[1], int(n[1])),
                "conv_umma": (ms[2], fl[2], int(n[2]))}

    def profile_table(self, max_rows: int = 256):
        """Per-shape rows (cls, Cin, N, K, dil, MT, launches, ms, flops) of the records profile_read() consumed."""
        W = _lib.PROFILE_COLS
        buf = (C.c_double * (W * max_rows))()
        n = self.lib.pg_profile_table(self._h, buf, max_rows)
        if n < 0:
            raise RuntimeError(self.lib.pg_last_error().decode())
        return [tuple(buf[W * i + k] for k in range(W)) for i in range(n)]

    def _on_capturable_stream(self, fn):
        """Run fn(stream_handle) on the caller's current stream -- or, when that is the legacy default
        stream (which CUDA cannot capture, so pg_infer would never replay a graph), on a private side
        stream ordered after and before it.  This is what lets the plain drop-in call
        ``net_g.infer(...)`` of rvc/infer/pipeline.py:275 use the graph path."""
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != 0:
            return fn(C.c_void_p(cur.cuda_stream))
        side = self.__dict__.get("_side")
        if side is None:
            with torch.cuda.device(self.device):
                side = self.__dict__["_side"] = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        out = fn(C.c_void_p(side.cuda_stream))
        cur.wait_stream(side)
        return out

    def padded_frames(self, T: int) -> int:
        """the launch-shape bucket a T-frame call runs at (one CUDA graph per (B, bucket))"""
        return int(self.lib.pg_padded_frames(self._h, int(T)))

    def set_graph_cache(self, max_graphs: int):
        _lib.check(self.lib.pg_set_graph_cache(self._h, int(max_graphs)), "pg_set_graph_cache")

    def set_decoder_sms(self, sms: int):
        """cap the CTAs of the decoder's persistent kernels (0 = one per SM); see pg_set_decoder_sms"""
        _lib.check(self.lib.pg_set_decoder_sms(self._h, int(sms)), "pg_set_decoder_sms")

    def infer(self, phone, lengths, pitch, f0, sid, eps_zp=None, eps_src=None, seed: int = 0,
              want_aux: bool = True, wave_out=None):
        """Time-major tensors on this engine's GPU.  Returns (wave [B][L], aux [4][B][T][C] | None).
        `wave_out` (optional) receives the waveform instead of a fresh tensor.  The call is staged
        (include/polgen_rvc.h: pg_infer), so the inputs / `wave_out` may also be pinned CPU tensors."""
        B, T, _ = phone.shape
        dev = self._dev()
        wave = wave_out if wave_out is not None else torch.empty(B, T * self.cfg.upp, device=dev, dtype=torch.float32)
        aux = torch.empty(4, B, T, self.cfg.inter_channels, device=dev, dtype=torch.float32) if want_aux else None
        self._on_capturable_stream(lambda st: _lib.check(
            self.lib.pg_infer(self._h, st, B, T, self._ptr(phone), self._ptr(lengths),
                              self._ptr(pitch), self._ptr(f0), self._ptr(sid), self._ptr(eps_zp),
                              self._ptr(eps_src), C.c_uint64(seed & (2 ** 64 - 1)), self._ptr(wave),
                              self._ptr(aux)), "pg_infer"))
        return wave, aux

    def infer_segments(self, segs, seed: int = 0, trim: int = 0, want_aux: bool = False):
        """The silence-split segments of a clip as ONE ragged call (pg_infer_segments; the loop of
        rvc/infer/pipeline.py:381-447).  `segs`: list of dicts with phone [T][D] f32, pitch [T] i64,
        f0 [T] f32, sid (int) and optionally eps_zp [T][C], eps_src [T*upp], wave (out buffer of
        T*upp - 2*trim f32); tensors on this GPU or pinned CPU.  Returns (waves, auxes): every wave has
        `trim` samples dropped at both ends (t_pad_tgt, pipeline.py:397)."""
        n = len(segs)
        arr = (_lib.PgSegment * n)()
        dev = self._dev()
        waves, auxes, keep = [], [], []
        for i, g in enumerate(segs):
            phone, pitch, f0 = g["phone"], g["pitch"], g["f0"]
            T = phone.shape[-2]
            for t in (phone, pitch, f0):
                if not t.is_contiguous():
                    raise ValueError("segment tensors must be contiguous")
            wave = g.get("wave")
            if wave is None:
                wave = torch.empty(T * self.cfg.upp - 2 * trim, device=dev, dtype=torch.float32)
            aux = torch.empty(4, T, self.cfg.inter_channels, device=dev, dtype=torch.float32) if want_aux else None
            arr[i].T = T
            arr[i].trim = trim
            arr[i].sid = int(g.get("sid", 0))
            arr[i].phone = phone.data_ptr()
            arr[i].pitch = pitch.data_ptr()
            arr[i].f0 = f0.data_ptr()
            ez, es = g.get("eps_zp"), g.get("eps_src")
            arr[i].eps_zp = 0 if ez is None else ez.data_ptr()
            arr[i].eps_src = 0 if es is None else es.data_ptr()
            arr[i].wave = wave.data_ptr()
            arr[i].aux = 0 if aux is None else aux.data_ptr()
            waves.append(wave)
            auxes.append(aux)
            keep.append((phone, pitch, f0, ez, es))
        self._on_capturable_stream(lambda st: _lib.check(
            self.lib.pg_infer_segments(self._h, st, n, arr, C.c_uint64(seed & (2 ** 64 - 1))), "pg_infer_segments"))
        # the copies are asynchronous: the inputs of the last few calls stay referenced
        self._keep_alive = (getattr(self, "_keep_alive", []) + [keep])[-4:]
        return waves, auxes

    def text_encoder(self, phone, lengths, pitch):
        B, T, _ = phone.shape
        m_p = torch.empty(B, T, self.cfg.inter_channels, device=self._dev(), dtype=torch.float32)
        logs_p = torch.empty_like(m_p)
        _lib.check(self.lib.pg_text_encoder(self._h, self._stream(), B, T, self._ptr(phone),
                                            self._ptr(lengths), self._ptr(pitch), self._ptr(m_p),
                                            self._ptr(logs_p)), "pg_text_encoder")
        return m_p, logs_p

    def flow_reverse(self, z_p, lengths, sid):
        B, T, _ = z_p.shape
        z = torch.empty_like(z_p)
        _lib.check(self.lib.pg_flow_reverse(self._h, self._stream(), B, T, self._ptr(z_p),
                                            self._ptr(lengths), self._ptr(sid), self._ptr(z)),
                   "pg_flow_reverse")
        return z

    def source(self, f0, eps_src=None, seed: int = 0, want_sine: bool = False):
        B, T = f0.shape
        src = torch.empty(B, T * self.cfg.upp, device=self._dev(), dtype=torch.float32)
        sine = torch.empty_like(src) if want_sine else None
        _lib.check(self.lib.pg_source(self._h, self._stream(), B, T, self._ptr(f0), self._ptr(eps_src),
                                      C.c_uint64(seed & (2 ** 64 - 1)), self._ptr(src), self._ptr(sine)),
                   "pg_source")
        return src, sine

    def generator(self, z, source, sid):
        B, T, _ = z.shape
        wave = torch.empty(B, T * self.cfg.upp, device=self._dev(), dtype=torch.float32)
        _lib.check(self.lib.pg_generator(self._h, self._stream(), B, T, self._ptr(z), self._ptr(source),
                                         self._ptr(sid), self._ptr(wave)), "pg_generator")
        return wave

    def fetch_tap(self, name: str) -> torch.Tensor:
        shape = (C.c_int64 * 3)()
        n = self.lib.pg_debug_fetch(self._h, self._stream(), name.encode(), None, 0, shape)
        if n < 0:
            _lib.check(int(n), f"pg_debug_fetch({name})")
        out = torch.empty(tuple(int(s) for s in shape), device=self._dev(), dtype=torch.float32)
        n = self.lib.pg_debug_fetch(self._h, self._stream(), name.encode(), self._ptr(out), out.numel(), shape)
        if n < 0:
            _lib.check(int(n), f"pg_debug_fetch({name})")
        return out


# --------------------------------------------------------------------------
# parameter containers (names == the reference's state-dict keys)
# --------------------------------------------------------------------------
def _normal_(m: nn.Module, std: float = 0.01):
    # commons.py:7-10 init_weights
    m.weight.data.normal_(0.0, std)


class _Holder(nn.Module):
    def forward(self, *a, **k):   # pragma: no cover
        raise RuntimeError("parameter container: compute runs in the CUDA library")


def _attn(H: int, heads: int, window: int) -> nn.Module:
    m = _Holder()
    for n in "qkvo":
        setattr(m, f"conv_{n}", nn.Conv1d(H, H, 1))
    d = H // heads
    m.emb_rel_k = nn.Parameter(torch.randn(1, 2 * window + 1, d) * d ** -0.5)
    m.emb_rel_v = nn.Parameter(torch.randn(1, 2 * window + 1, d) * d ** -0.5)
    for n in "qkv":
        nn.init.xavier_uniform_(getattr(m, f"conv_{n}").weight)
    return m


def _ln(H: int) -> nn.Module:
    m = _Holder()
    m.gamma = nn.Parameter(torch.ones(H))
    m.beta = nn.Parameter(torch.zeros(H))
    return m


def _text_encoder(cfg: SynthConfig) -> nn.Module:
    H, F = cfg.hidden_channels, cfg.filter_channels
    enc = _Holder()
    enc.attn_layers = nn.ModuleList(_attn(H, cfg.n_heads, cfg.attn_window) for _ in range(cfg.n_layers))
    enc.norm_layers_1 = nn.ModuleList(_ln(H) for _ in range(cfg.n_layers))
    ffn = []
    for _ in range(cfg.n_layers):
        f = _Holder()
        f.conv_1 = nn.Conv1d(H, F, cfg.kernel_size)
        f.conv_2 = nn.Conv1d(F, H, cfg.kernel_size)
        ffn.append(f)
    enc.ffn_layers = nn.ModuleList(ffn)
    enc.norm_layers_2 = nn.ModuleList(_ln(H) for _ in range(cfg.n_layers))
    te = _Holder()
    te.emb_phone = nn.Linear(cfg.input_dim, H)
    te.emb_pitch = nn.Embedding(256, H)
    te.encoder = enc
    te.proj = nn.Conv1d(H, 2 * cfg.inter_channels, 1)
    return te


def _flow(cfg: SynthConfig) -> nn.Module:
    H, half = cfg.hidden_channels, cfg.inter_channels // 2
    nl, kw = cfg.flow_wn_layers, cfg.flow_wn_kernel
    mods = []
    for _ in range(cfg.flow_n_flows):
        wn = _Holder()
        # registration order follows modules.py:28-56 so state_dict() key order matches too
        wn.in_layers = nn.ModuleList()
        wn.res_skip_layers = nn.ModuleList()
        wn.cond_layer = weight_norm(nn.Conv1d(cfg.gin_channels, 2 * H * nl, 1), name="weight")
        for l in range(nl):
            wn.in_layers.append(weight_norm(nn.Conv1d(H, 2 * H, kw, padding=(kw - 1) // 2), name="weight"))
            wn.res_skip_layers.append(
                weight_norm(nn.Conv1d(H, H if l == nl - 1 else 2 * H, 1), name="weight"))
        rcl = _Holder()
        rcl.pre = nn.Conv1d(half, H, 1)
        rcl.enc = wn
        rcl.post = nn.Conv1d(H, half, 1)
        rcl.post.weight.data.zero_()     # residuals.py:207-208
        rcl.post.bias.data.zero_()
        mods += [rcl, _Holder()]          # coupling layer, Flip (no parameters)
    f = _Holder()
    f.flows = nn.ModuleList(mods)
    return f


def _decoder(cfg: SynthConfig) -> nn.Module:
    dec = _Holder()
    dec.m_source = _Holder()
    dec.m_source.l_linear = nn.Linear(1, 1)
    c0 = cfg.upsample_initial_channel
    dec.conv_pre = nn.Conv1d(cfg.inter_channels, c0, 7, 1, padding=3)
    ups, noise, blocks = [], [], []
    cin = c0
    strides = cfg.noise_strides()
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        cout = cin // 2
        up = nn.ConvTranspose1d(cin, cout, k, u, padding=(k - u) // 2)
        _normal_(up)
        ups.append(weight_norm(up))
        s = strides[i]
        noise.append(nn.Conv1d(1, cout, kernel_size=(2 * s if s > 1 else 1), stride=s,
                               padding=(s // 2 if s > 1 else 0)))
        for ksz, dils in zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes):
            rb = _Holder()
            c1, c2 = [], []
            for d in dils:
                a = nn.Conv1d(cout, cout, ksz, 1, dilation=d, padding=(ksz * d - d) // 2)
                b = nn.Conv1d(cout, cout, ksz, 1, dilation=1, padding=(ksz - 1) // 2)
                _normal_(a)
                _normal_(b)
                c1.append(weight_norm(a))
                c2.append(weight_norm(b))
            rb.convs1 = nn.ModuleList(c1)
            rb.convs2 = nn.ModuleList(c2)
            blocks.append(rb)
        cin = cout
    dec.ups = nn.ModuleList(ups)
    dec.noise_convs = nn.ModuleList(noise)
    dec.resblocks = nn.ModuleList(blocks)
    dec.conv_post = nn.Conv1d(cin, 1, 7, 1, padding=3, bias=False)
    dec.cond = nn.Conv1d(cfg.gin_channels, c0, 1)
    dec.upp = cfg.upp
    return dec


class Synthesizer(nn.Module):
    """B200-native drop-in for ``rvc.lib.algorithm.synthesizers.Synthesizer``."""

    def __init__(self, spec_channels, segment_size, inter_channels, hidden_channels, filter_channels,
                 n_heads, n_layers, kernel_size, p_dropout, resblock, resblock_kernel_sizes,
                 resblock_dilation_sizes, upsample_rates, upsample_initial_channel,
                 upsample_kernel_sizes, spk_embed_dim, gin_channels, sr, use_f0, input_dim=768,
                 **kwargs):
        super().__init__()
        if not use_f0:
            # the reference's non-F0 Generator has no forward (generators.py:57) and
            # pipeline.py:282 passes sid in the pitch slot: nothing to be compatible with
            raise NotImplementedError("only the F0 (NSF) synthesizer is supported")
        if "is_half" not in kwargs:
            raise KeyError("is_half")   # synthesizers.py:81 reads kwargs["is_half"]
        if str(resblock) != "1":
            raise NotImplementedError("only resblock='1' (ResBlock1) is supported")
        self.is_half = bool(kwargs["is_half"])
        args = [spec_channels, segment_size, inter_channels, hidden_channels, filter_channels, n_heads,
                n_layers, kernel_size, p_dropout, resblock, resblock_kernel_sizes,
                resblock_dilation_sizes, upsample_rates, upsample_initial_channel,
                upsample_kernel_sizes, spk_embed_dim, gin_channels, sr]
        self.cfg = config_from_ctor(args, input_dim=input_dim)
        # the attributes the reference exposes
        self.spec_channels = spec_channels
        self.inter_channels = inter_channels
        self.hidden_channels = hidden_channels
        self.filter_channels = filter_channels
        self.n_heads = n_heads
        self.n_layers = n_layers
        self.kernel_size = kernel_size
        self.p_dropout = float(p_dropout)
        self.resblock = resblock
        self.resblock_kernel_sizes = resblock_kernel_sizes
        self.resblock_dilation_sizes = resblock_dilation_sizes
        self.upsample_rates = upsample_rates
        self.upsample_initial_channel = upsample_initial_channel
        self.upsample_kernel_sizes = upsample_kernel_sizes
        self.segment_size = segment_size
        self.gin_channels = gin_channels
        self.spk_embed_dim = spk_embed_dim
        self.use_f0 = use_f0

        self.enc_p = _text_encoder(self.cfg)
        self.dec = _decoder(self.cfg)
        self.enc_q = _Holder()            # training-only posterior encoder; callers `del` it (infer.py:99)
        self.flow = _flow(self.cfg)
        self.emb_g = nn.Embedding(self.spk_embed_dim, gin_channels)
        self._engine: Optional[Engine] = None
        self._engine_flags = 0

    # -- weight lifecycle: any change to the parameters drops the device copy --
    def _invalidate(self):
        eng = self.__dict__.get("_engine")
        if eng is not None:
            eng.close()
        self.__dict__["_engine"] = None

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        self._invalidate()
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    def _apply(self, fn, recurse=True):
        self._invalidate()
        return super()._apply(fn, recurse)

    def remove_weight_norm(self):
        """The reference's version raises on the inference path; here the fold
        happens when the weights are uploaded, so this is a no-op."""
        return self

    def _device_index(self) -> int:
        p = self.emb_g.weight
        if p.device.type != "cuda":
            raise RuntimeError("polgen-rvc_b200 Synthesizer runs on CUDA (sm_100a) only: move the module "
                               "with .to('cuda') -- there is no CPU path")
        return p.device.index if p.device.index is not None else torch.cuda.current_device()

    def engine(self) -> Engine:
        if self._engine is None:
            dev = self._device_index()
            self.__dict__["_engine"] = Engine(self.cfg, fold_state_dict(self.state_dict()), dev,
                                              self._engine_flags)
        return self._engine

    @torch.jit.ignore
    def forward(self, *args, **kwargs):
        raise NotImplementedError("training forward is out of scope (synthesizers.py:136-160)")

    @torch.no_grad()
    def infer(self, phone: torch.Tensor, phone_lengths: torch.Tensor, pitch: Optional[torch.Tensor] = None,
              nsff0: Optional[torch.Tensor] = None, sid: torch.Tensor = None,
              rate: Optional[torch.Tensor] = None, *, eps_zp: Optional[torch.Tensor] = None,
              eps_src: Optional[torch.Tensor] = None):
        """``synthesizers.py:162-188``.  ``eps_zp`` (B, C, T) / ``eps_src`` (B, T*upp, 1) replace the two
        in-path ``randn_like`` draws (parity tests); otherwise noise is drawn on device from Philox
        seeded by torch's CPU generator."""
        if pitch is None or nsff0 is None or sid is None:
            raise ValueError("pitch, nsff0 and sid are required (F0 synthesizer)")
        if not phone.is_cuda:
            raise RuntimeError("polgen-rvc_b200 has no CPU path: pass CUDA tensors")
        eng = self.engine()
        out_dtype = phone.dtype
        dev = phone.device
        B, T, _ = phone.shape
        phone32 = phone.to(torch.float32).contiguous()
        lengths = phone_lengths.to(dev, torch.int64).contiguous()
        pitch64 = pitch.to(dev, torch.int64).contiguous()
        f0 = nsff0.to(dev, torch.float32).contiguous()
        sid64 = sid.to(dev, torch.int64).reshape(-1).contiguous()
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        C_ = self.cfg.inter_channels
        if rate is None:
            ez = None if eps_zp is None else eps_zp.to(dev, torch.float32).transpose(1, 2).contiguous()
            es = None if eps_src is None else eps_src.to(dev, torch.float32).reshape(B, -1).contiguous()
            wave, aux = eng.infer(phone32, lengths, pitch64, f0, sid64, ez, es, seed)
            z, z_p, m_p, logs_p = (aux[i].transpose(1, 2) for i in range(4))
            Tm = T
        else:
            # synthesizers.py:175-181: drop the head of the latent before the flow
            assert isinstance(rate, torch.Tensor)
            m_t, logs_t = eng.text_encoder(phone32, lengths, pitch64)
            mask_t = (torch.arange(T, device=dev)[None, :] < lengths[:, None]).to(torch.float32)[:, :, None]
            e = torch.randn_like(m_t) if eps_zp is None else eps_zp.to(dev, torch.float32).transpose(1, 2)
            zp_t = (m_t + torch.exp(logs_t) * e * 0.66666) * mask_t
            head = int(T * (1.0 - rate.item()))
            zp_t = zp_t[:, head:].contiguous()
            f0 = f0[:, head:].contiguous()
            Tm = T - head
            # the flow/decoder mask is the sliced x_mask: rows keep length - head valid frames
            len2 = (lengths - head).clamp(min=0)
            z_t = eng.flow_reverse(zp_t, len2, sid64)
            es = None if eps_src is None else eps_src.to(dev, torch.float32).reshape(B, -1).contiguous()
            src, _ = eng.source(f0, es, seed)
            m2 = (torch.arange(Tm, device=dev)[None, :] < len2[:, None]).to(torch.float32)[:, :, None]
            wave = eng.generator((z_t * m2).contiguous(), src, sid64)
            z, z_p = z_t.transpose(1, 2), zp_t.transpose(1, 2)
            m_p, logs_p = m_t.transpose(1, 2), logs_t.transpose(1, 2)
            lengths = len2
        x_mask = (torch.arange(Tm, device=dev)[None, :] < lengths[:, None]).to(out_dtype)[:, None, :]
        o = wave[:, None, :].to(out_dtype)
        return o, x_mask, (z.to(out_dtype), z_p.to(out_dtype), m_p.to(out_dtype), logs_p.to(out_dtype))


# upstream-RVC class names (BASELINE.json north_star): thin aliases that pin input_dim
class SynthesizerTrnMs768NSFsid(Synthesizer):
    def __init__(self, *args, **kwargs):
        kwargs.setdefault("use_f0", 1)
        kwargs["input_dim"] = 768
        super().__init__(*args, **kwargs)


class SynthesizerTrnMs256NSFsid(Synthesizer):
    def __init__(self, *args, **kwargs):
        kwargs.setdefault("use_f0", 1)
        kwargs["input_dim"] = 256
        super().__init__(*args, **kwargs)


# --------------------------------------------------------------------------
# Segment scheduler: the reference decodes the silence-split segments of
# a clip one after another as B=1 infer calls (rvc/infer/pipeline.py:381-447).  Segments are
# independent, so this host-side scheduler
#   * packs them into ragged batches (pg_infer_segments: every layer treats a row's own length as
#     its hard end, so a row equals its stand-alone B=1 decode whatever it is batched with),
#     which divides the number of latency-bound TextEncoder / flow launches per clip, and
#   * deals the batches over `lanes` engines (handle + workspace + CUDA stream each) on one GPU, so
#     the TextEncoder / flow phase of one batch overlaps the SM-filling decoder phase of another.
# --------------------------------------------------------------------------
class SegmentScheduler:
    def __init__(self, cfg: SynthConfig, weights: Dict[str, torch.Tensor], device: int = 0, lanes: int = 2,
                 flags: int = 0, max_batch: int = 64, max_batch_frames: int = 36000, decoder_sms: int = 0):
        self.cfg = cfg
        self.device = int(device)
        self.max_batch = int(max_batch)
        self.max_batch_frames = int(max_batch_frames)     # ~430 KB of workspace per frame
        self.engines = [Engine(cfg, weights, device, flags) for _ in range(max(1, lanes))]
        if decoder_sms > 0 and len(self.engines) > 1:
            # leave a few SMs outside the decoder's persistent grids: the short TextEncoder / flow kernels of
            # one lane's next clip then run beside the other lane's decoder instead of queueing behind it
            for e in self.engines:
                e.set_decoder_sms(decoder_sms)
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream(device=self.device) for _ in self.engines]
        self._next_lane = 0

    def close(self):
        for e in self.engines:
            e.close()

    def plan_batches(self, lengths):
        """Indices grouped into ragged batches: longest first, a batch grows while it stays within
        max_batch rows and max_batch_frames padded frames (rows x bucket of its longest row)."""
        order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
        batches, cur, cur_T = [], [], 0
        for i in order:
            T = self.engines[0].padded_frames(int(lengths[i]))
            top = max(cur_T, T)
            if cur and (len(cur) + 1 > self.max_batch or (len(cur) + 1) * top > self.max_batch_frames):
                batches.append(cur)
                cur, top = [], T
            cur.append(i)
            cur_T = top
        if cur:
            batches.append(cur)
        return batches

    @staticmethod
    def _segment_dict(seg):
        """(phone, lengths, pitch, f0, sid) tensors of one B=1 segment (the reference's infer arguments) or a
        ready dict -> the row description Engine.infer_segments takes"""
        if isinstance(seg, dict):
            return seg
        phone, _, pitch, f0, sid = seg
        sid = int(sid.reshape(-1)[0]) if torch.is_tensor(sid) else int(sid)   # a CUDA sid costs a sync: pass CPU
        return {"phone": phone.reshape(-1, phone.shape[-1]), "pitch": pitch.reshape(-1), "f0": f0.reshape(-1),
                "sid": sid}

    def decode(self, segments, seed: Optional[int] = None, host_out=None, join=True, out=None, trim: int = 0):
        """segments: list of (phone, lengths, pitch, f0, sid) per segment (B = 1 each; CUDA tensors on
        this GPU or pinned CPU tensors -- sid a CPU tensor or int).  Returns the list of waveforms
        [1][T*upp - 2*trim]: fresh CUDA tensors, or the caller's buffers when given -- `host_out` (pinned
        CPU tensors, filled by async copies; the host is synchronised before returning) or `out`
        (preallocated CUDA tensors: a steady stream of clips then makes no allocator calls at all; they
        may be views into one clip-long buffer, see `trim`).  `trim` drops that many samples at both ends
        of every waveform (t_pad_tgt, pipeline.py:397).  `seed` None draws a fresh base seed from torch's
        CPU generator (as the reference draws fresh noise per call); batch k of the call uses seed + k
        and row b of a batch Philox(seed_k + b*0x9E37...).  The caller's current stream waits for every
        lane before this returns control to it.  join=False (streaming use: clip after clip) skips both
        joins, so the lanes run on into the next call and their phases interleave freely; call join()
        before touching the results.  Inputs must then already be complete on the device (or pinned)."""
        dev = torch.device("cuda", self.device)
        cur = torch.cuda.current_stream(self.device)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        ev0 = None
        if join:
            ev0 = torch.cuda.Event()
            ev0.record(cur)
        rows = [self._segment_dict(s) for s in segments]
        lengths = [r["phone"].shape[0] for r in rows]
        dst = host_out if host_out is not None else out
        outs = [None] * len(rows)
        self.last_launches = 0
        self.last_batches = self.plan_batches(lengths)
        for k, idx in enumerate(self.last_batches):
            lane = self._next_lane
            self._next_lane = (lane + 1) % len(self.engines)
            st = self.streams[lane]
            if ev0 is not None:
                st.wait_event(ev0)
            with torch.cuda.stream(st):
                batch = []
                for i in idx:
                    r = dict(rows[i])
                    for key in ("phone", "pitch", "f0"):
                        t = r[key]
                        if not (t.is_cuda or t.is_pinned()):
                            t = t.to(dev, non_blocking=True)      # pageable host memory: through torch
                        r[key] = t.contiguous()
                    n_out = lengths[i] * self.cfg.upp - 2 * trim
                    if dst is not None:
                        w = dst[i].reshape(-1)
                        if w.numel() != n_out or not w.is_contiguous():
                            raise ValueError(f"output buffer {i} must hold {n_out} contiguous samples")
                        r["wave"] = w
                    batch.append(r)
                waves, _ = self.engines[lane].infer_segments(batch, seed + k, trim=trim)
                self.last_launches += self.engines[lane].launch_count()
                for i, w in zip(idx, waves):
                    outs[i] = dst[i] if dst is not None else w.reshape(1, -1)
                    if dst is None:
                        w.record_stream(cur)
        if join:
            self.join(host_sync=host_out is not None)
        return outs

    def join(self, host_sync: bool = False):
        """make the caller's current stream wait for every lane (and optionally the host too)"""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            ev = torch.cuda.Event()
            ev.record(st)
            cur.wait_event(ev)
        if host_sync:
            cur.synchronize()

    def launch_count(self) -> int:
        """kernels launched by the last decode() call"""
        return getattr(self, "last_launches", 0)
