"""Hyper-parameter tables and the state-dict inventory of the RVC synthesizer.

The reference repo ships no JSON configs: the 18 positional constructor values
arrive inside the checkpoint (``cpt["config"]``, reference
``rvc/infer/infer.py:86-97``).  The tables below are the upstream RVC v1/v2
training configs that BASELINE.json names.

``state_dict_shapes`` enumerates every tensor the reference ``Synthesizer``
(``rvc/lib/algorithm/synthesizers.py:13-112``, ``enc_q`` stripped as
``infer.py:99`` does) owns, in the legacy checkpoint spelling
(``weight_g``/``weight_v``), and ``synth_weights`` fills them with a
deterministic, construction-order-independent pseudo-random initialisation of
the same distributions the reference uses.  It is what
tests, ``bench.py`` and the golden-vector generator share, so the GPU box can
rebuild the exact weights the fixtures were made with, without the reference.
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch


@dataclass(frozen=True)
class SynthConfig:
    name: str
    spec_channels: int
    segment_size: int
    inter_channels: int
    hidden_channels: int
    filter_channels: int
    n_heads: int
    n_layers: int
    kernel_size: int
    p_dropout: float
    resblock: str
    resblock_kernel_sizes: Tuple[int, ...]
    resblock_dilation_sizes: Tuple[Tuple[int, ...], ...]
    upsample_rates: Tuple[int, ...]
    upsample_initial_channel: int
    upsample_kernel_sizes: Tuple[int, ...]
    spk_embed_dim: int
    gin_channels: int
    sr: int
    input_dim: int = 768
    # fixed by the reference code, not by the checkpoint
    flow_n_flows: int = 4          # residuals.py:116
    flow_wn_layers: int = 3        # synthesizers.py:104
    flow_wn_kernel: int = 5        # synthesizers.py:104
    attn_window: int = 10          # encoders.py:22

    @property
    def upp(self) -> int:
        return math.prod(self.upsample_rates)

    def ctor_args(self) -> list:
        """The 18 positional values of ``cpt["config"]`` (infer.py:92)."""
        return [
            self.spec_channels, self.segment_size, self.inter_channels,
            self.hidden_channels, self.filter_channels, self.n_heads,
            self.n_layers, self.kernel_size, self.p_dropout, self.resblock,
            [int(k) for k in self.resblock_kernel_sizes],
            [list(d) for d in self.resblock_dilation_sizes],
            list(self.upsample_rates), self.upsample_initial_channel,
            list(self.upsample_kernel_sizes), self.spk_embed_dim,
            self.gin_channels, self.sr,
        ]

    def stage_channels(self) -> List[int]:
        return [self.upsample_initial_channel // (2 ** (i + 1))
                for i in range(len(self.upsample_rates))]

    def noise_strides(self) -> List[int]:
        r = self.upsample_rates
        return [math.prod(r[i + 1:]) if i + 1 < len(r) else 1 for i in range(len(r))]

    def generator_flops_per_frame(self) -> float:
        """2*MAC count of the GeneratorNSF conv layers per input frame
        (110.15 GFLOP per audio-second at 48k = 100 frames)."""
        c0 = self.upsample_initial_channel
        fl = 2.0 * self.inter_channels * c0 * 7          # conv_pre
        length = 1
        cin = c0
        for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
            cout = cin // 2
            fl += 2.0 * length * cin * cout * k          # transposed conv: per INPUT sample
            length *= u
            s = self.noise_strides()[i]
            fl += 2.0 * length * cout * (2 * s if s > 1 else 1)
            for kk, dd in zip(self.resblock_kernel_sizes, self.resblock_dilation_sizes):
                fl += 2.0 * length * cout * cout * kk * 2 * len(dd)
            cin = cout
        fl += 2.0 * length * cin * 7                      # conv_post
        return fl


def _mk(name, sr, rates, ksz, spec, input_dim) -> SynthConfig:
    return SynthConfig(
        name=name, spec_channels=spec, segment_size=32, inter_channels=192,
        hidden_channels=192, filter_channels=768, n_heads=2, n_layers=6,
        kernel_size=3, p_dropout=0.0, resblock="1",
        resblock_kernel_sizes=(3, 7, 11),
        resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)),
        upsample_rates=tuple(rates), upsample_initial_channel=512,
        upsample_kernel_sizes=tuple(ksz), spk_embed_dim=109, gin_channels=256,
        sr=sr, input_dim=input_dim)


CONFIGS: Dict[str, SynthConfig] = {
    "v2-48k": _mk("v2-48k", 48000, (12, 10, 2, 2), (24, 20, 4, 4), 1025, 768),
    "v2-40k": _mk("v2-40k", 40000, (10, 10, 2, 2), (16, 16, 4, 4), 1025, 768),
    "v2-32k": _mk("v2-32k", 32000, (10, 8, 2, 2), (20, 16, 4, 4), 513, 768),
    "v1-40k": _mk("v1-40k", 40000, (10, 10, 2, 2), (16, 16, 4, 4), 1025, 256),
}


def config_from_ctor(args, input_dim=768, name="custom") -> SynthConfig:
    """Build a SynthConfig from the reference constructor's 18 positionals."""
    (spec, seg, inter, hidden, filt, heads, layers, ksz, pdrop, resblock, rks,
     rds, ur, uic, uks, spk, gin, sr) = args
    return SynthConfig(
        name=name, spec_channels=int(spec), segment_size=int(seg),
        inter_channels=int(inter), hidden_channels=int(hidden),
        filter_channels=int(filt), n_heads=int(heads), n_layers=int(layers),
        kernel_size=int(ksz), p_dropout=float(pdrop), resblock=str(resblock),
        resblock_kernel_sizes=tuple(int(k) for k in rks),
        resblock_dilation_sizes=tuple(tuple(int(x) for x in d) for d in rds),
        upsample_rates=tuple(int(u) for u in ur), upsample_initial_channel=int(uic),
        upsample_kernel_sizes=tuple(int(k) for k in uks), spk_embed_dim=int(spk),
        gin_channels=int(gin), sr=int(sr), input_dim=int(input_dim))


# --------------------------------------------------------------------------
# state-dict inventory
# --------------------------------------------------------------------------
# kinds: how synth_weights draws the tensor
#   "conv"   kaiming-uniform(a=sqrt5) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
#   "xavier" xavier-uniform (attentions.py:55-57)
#   "bias:<fan_in>" U(-1/sqrt(fan_in), 1/sqrt(fan_in))
#   "n01"    N(0, 1) (nn.Embedding)
#   "rel"    N(0, k_channels^-0.5)  (attentions.py:45-53)
#   "v001"   N(0, 0.01) weight_v of init_weights'd convs (commons.py:7-10)
#   "g:<key of v>:<dim>" weight_g = ||v|| over all dims except <dim>, jittered
#   "ones"/"zeros" LayerNorm gamma/beta
#   "post"   coupling `post` conv: zeros in the reference (residuals.py:207-208)


def state_dict_shapes(cfg: SynthConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    out: List[Tuple[str, Tuple[int, ...], str]] = []
    H, F, inter = cfg.hidden_channels, cfg.filter_channels, cfg.inter_channels
    kc = H // cfg.n_heads
    out.append(("enc_p.emb_phone.weight", (H, cfg.input_dim), "conv"))
    out.append(("enc_p.emb_phone.bias", (H,), f"bias:{cfg.input_dim}"))
    out.append(("enc_p.emb_pitch.weight", (256, H), "n01"))
    for i in range(cfg.n_layers):
        p = f"enc_p.encoder.attn_layers.{i}."
        out.append((p + "emb_rel_k", (1, 2 * cfg.attn_window + 1, kc), "rel"))
        out.append((p + "emb_rel_v", (1, 2 * cfg.attn_window + 1, kc), "rel"))
        for n in "qkvo":
            out.append((p + f"conv_{n}.weight", (H, H, 1), "conv" if n == "o" else "xavier"))
            out.append((p + f"conv_{n}.bias", (H,), f"bias:{H}"))
    for i in range(cfg.n_layers):
        out.append((f"enc_p.encoder.norm_layers_1.{i}.gamma", (H,), "ones"))
        out.append((f"enc_p.encoder.norm_layers_1.{i}.beta", (H,), "zeros"))
    for i in range(cfg.n_layers):
        p = f"enc_p.encoder.ffn_layers.{i}."
        out.append((p + "conv_1.weight", (F, H, cfg.kernel_size), "conv"))
        out.append((p + "conv_1.bias", (F,), f"bias:{H * cfg.kernel_size}"))
        out.append((p + "conv_2.weight", (H, F, cfg.kernel_size), "conv"))
        out.append((p + "conv_2.bias", (H,), f"bias:{F * cfg.kernel_size}"))
    for i in range(cfg.n_layers):
        out.append((f"enc_p.encoder.norm_layers_2.{i}.gamma", (H,), "ones"))
        out.append((f"enc_p.encoder.norm_layers_2.{i}.beta", (H,), "zeros"))
    out.append(("enc_p.proj.weight", (2 * inter, H, 1), "conv"))
    out.append(("enc_p.proj.bias", (2 * inter,), f"bias:{H}"))

    # decoder (nsf.py:43-118)
    c0 = cfg.upsample_initial_channel
    out.append(("dec.m_source.l_linear.weight", (1, 1), "conv"))
    out.append(("dec.m_source.l_linear.bias", (1,), "bias:1"))
    out.append(("dec.conv_pre.weight", (c0, inter, 7), "conv"))
    out.append(("dec.conv_pre.bias", (c0,), f"bias:{inter * 7}"))
    chans = cfg.stage_channels()
    strides = cfg.noise_strides()
    cin = c0
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        cout = chans[i]
        # ConvTranspose1d: fan_in as torch computes it = weight.size(1)*k = cout*k
        out.append((f"dec.ups.{i}.bias", (cout,), f"bias:{cout * k}"))
        out.append((f"dec.ups.{i}.weight_g", (cin, 1, 1), f"g:dec.ups.{i}.weight_v:0"))
        out.append((f"dec.ups.{i}.weight_v", (cin, cout, k), "v001"))
        cin = cout
    for i, cout in enumerate(chans):
        s = strides[i]
        kn = 2 * s if s > 1 else 1
        out.append((f"dec.noise_convs.{i}.weight", (cout, 1, kn), "conv"))
        out.append((f"dec.noise_convs.{i}.bias", (cout,), f"bias:{kn}"))
    j = 0
    for i, c in enumerate(chans):
        for k, dil in zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes):
            for grp in ("convs1", "convs2"):
                for d in range(len(dil)):
                    p = f"dec.resblocks.{j}.{grp}.{d}."
                    out.append((p + "bias", (c,), f"bias:{c * k}"))
                    out.append((p + "weight_g", (c, 1, 1), f"g:{p}weight_v:0"))
                    out.append((p + "weight_v", (c, c, k), "v001"))
            j += 1
    out.append(("dec.conv_post.weight", (1, chans[-1], 7), "conv"))
    out.append(("dec.cond.weight", (c0, cfg.gin_channels, 1), "conv"))
    out.append(("dec.cond.bias", (c0,), f"bias:{cfg.gin_channels}"))

    # flow (residuals.py:109-232, modules.py:9-56)
    half = inter // 2
    nl, kw = cfg.flow_wn_layers, cfg.flow_wn_kernel
    for f in range(cfg.flow_n_flows):
        p = f"flow.flows.{2 * f}."
        out.append((p + "pre.weight", (H, half, 1), "conv"))
        out.append((p + "pre.bias", (H,), f"bias:{half}"))
        out.append((p + "enc.cond_layer.bias", (2 * H * nl,), f"bias:{cfg.gin_channels}"))
        out.append((p + "enc.cond_layer.weight_g", (2 * H * nl, 1, 1),
                    f"g:{p}enc.cond_layer.weight_v:0"))
        out.append((p + "enc.cond_layer.weight_v", (2 * H * nl, cfg.gin_channels, 1), "conv"))
        for l in range(nl):
            q = p + f"enc.in_layers.{l}."
            out.append((q + "bias", (2 * H,), f"bias:{H * kw}"))
            out.append((q + "weight_g", (2 * H, 1, 1), f"g:{q}weight_v:0"))
            out.append((q + "weight_v", (2 * H, H, kw), "conv"))
        for l in range(nl):
            q = p + f"enc.res_skip_layers.{l}."
            co = H if l == nl - 1 else 2 * H
            out.append((q + "bias", (co,), f"bias:{H}"))
            out.append((q + "weight_g", (co, 1, 1), f"g:{q}weight_v:0"))
            out.append((q + "weight_v", (co, H, 1), "conv"))
        out.append((p + "post.weight", (half, H, 1), "post"))
        out.append((p + "post.bias", (half,), "post"))
    out.append(("emb_g.weight", (cfg.spk_embed_dim, cfg.gin_channels), "n01"))
    return out


def _gen_for(key: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def synth_weights(cfg: SynthConfig, seed: int = 0, post_std: float = 0.01,
                  g_jitter: float = 0.2) -> Dict[str, torch.Tensor]:
    """Deterministic random-init state dict in the legacy checkpoint spelling.

    Each tensor is drawn from its own generator seeded by sha256(seed:key), so
    the result does not depend on construction order, on the torch module
    classes, or on the reference being importable.  ``post_std`` re-randomises
    the zero-initialised coupling ``post`` convs (otherwise the
    flow is an identity-with-flips and its kernels are not exercised);
    ``g_jitter`` scales ``weight_g`` away from ``||v||`` so the weight-norm
    fold is not a no-op.
    """
    sd: Dict[str, torch.Tensor] = {}
    shapes = state_dict_shapes(cfg)
    kc = cfg.hidden_channels // cfg.n_heads
    for key, shape, kind in shapes:
        if kind.startswith("g:"):
            continue
        gen = _gen_for(key, seed)
        if kind in ("conv", "xavier"):
            if len(shape) == 1:
                fan_in, fan_out = shape[0], shape[0]
            else:
                rf = math.prod(shape[2:]) if len(shape) > 2 else 1
                fan_in, fan_out = shape[1] * rf, shape[0] * rf
            if kind == "conv":
                b = 1.0 / math.sqrt(fan_in)
            else:
                b = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=gen) * 2 - 1) * b
        elif kind.startswith("bias:"):
            b = 1.0 / math.sqrt(int(kind.split(":")[1]))
            t = (torch.rand(shape, generator=gen) * 2 - 1) * b
        elif kind == "n01":
            t = torch.randn(shape, generator=gen)
        elif kind == "rel":
            t = torch.randn(shape, generator=gen) * kc ** -0.5
        elif kind == "v001":
            t = torch.randn(shape, generator=gen) * 0.01
        elif kind == "ones":
            t = torch.ones(shape) + 0.1 * torch.randn(shape, generator=gen)
        elif kind == "zeros":
            t = 0.1 * torch.randn(shape, generator=gen)
        elif kind == "post":
            t = torch.randn(shape, generator=gen) * post_std
        else:
            raise ValueError(kind)
        sd[key] = t.float().contiguous()
    for key, shape, kind in shapes:
        if not kind.startswith("g:"):
            continue
        _, vkey, dim = kind.split(":")
        v = sd[vkey]
        assert int(dim) == 0
        norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(shape)
        gen = _gen_for(key, seed)
        jit = 1.0 + g_jitter * (torch.rand(shape, generator=gen) * 2 - 1)
        sd[key] = (norm * jit).float().contiguous()
    return sd


def synth_inputs(cfg: SynthConfig, batch: int, frames: int, seed: int = 0):
    """Synthetic F0-swept inputs.

    phone ~ N(0,1); nsff0 = log sweep 80->800 Hz with one 0.5 s unvoiced gap and
    a per-row phase offset; pitch = coarse mel quantisation exactly as the
    reference caller does (``rvc/infer/pipeline.py:193-200``, f0_min 50,
    f0_max 1100 -> 1..255); phone_lengths = frames; sid = 0.
    """
    gen = torch.Generator(device="cpu")
    gen.manual_seed(1000 + seed)
    phone = torch.randn(batch, frames, cfg.input_dim, generator=gen)
    base = torch.linspace(0.0, 1.0, frames)
    f0 = torch.empty(batch, frames)
    for b in range(batch):
        ph = (base + 0.37 * b) % 1.0
        f0[b] = torch.exp(math.log(80.0) + ph * (math.log(800.0) - math.log(80.0)))
        if frames >= 8:
            gap = max(1, min(50, frames // 6))
            start = (frames // 3 + 7 * b) % max(1, frames - gap)
            f0[b, start:start + gap] = 0.0
    f0_mel_min = 1127.0 * math.log(1 + 50.0 / 700.0)
    f0_mel_max = 1127.0 * math.log(1 + 1100.0 / 700.0)
    mel = 1127.0 * torch.log(1 + f0 / 700.0)
    pos = mel > 0
    mel = torch.where(pos, (mel - f0_mel_min) * 254.0 / (f0_mel_max - f0_mel_min) + 1.0, mel)
    mel = torch.where(mel <= 1, torch.ones_like(mel), mel)
    mel = torch.where(mel > 255, torch.full_like(mel, 255.0), mel)
    pitch = torch.round(mel).long()   # np.rint in the reference
    lengths = torch.full((batch,), frames, dtype=torch.long)
    sid = torch.zeros(batch, dtype=torch.long)
    return phone, lengths, pitch, f0, sid


def synth_noise(cfg: SynthConfig, batch: int, frames: int, seed: int = 0):
    """The two in-path normal draws (synthesizers.py:174, generators.py:154) as
    caller-supplied tensors, in the reference's layouts: eps_zp (B, C, T),
    eps_src (B, T*upp, 1)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(2000 + seed)
    eps_zp = torch.randn(batch, cfg.inter_channels, frames, generator=gen)
    eps_src = torch.randn(batch, frames * cfg.upp, 1, generator=gen)
    return eps_zp, eps_src
