"""CPU oracle for the RVC synthesizer decode -- TEST INFRASTRUCTURE ONLY.

A from-scratch functional restatement (torch fp32 on CPU) of the reference hot
path ``Synthesizer.infer`` (reference ``rvc/lib/algorithm/synthesizers.py:162-188``)
operating on a plain state dict.  It exists to check the CUDA product path and
to time the CPU baseline; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.  The
product package never does.

Pinning: the reference ships no golden vectors, so this
oracle is pinned against outputs of the live reference module, imported
unmodified from ``/root/reference`` by ``oracle/make_golden.py`` and committed
under ``tests/golden/`` (inputs are regenerated deterministically from
``configs.synth_weights / synth_inputs / synth_noise``).
``tests/test_oracle_golden.py`` asserts the match; when ``/root/reference`` is
present it additionally compares against the live module.

Layouts follow the reference: activations are (B, C, T).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ---------------------------------------------------------------------------
# weight handling
# ---------------------------------------------------------------------------
def fold_weight_norm(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """w = g * v / ||v||, norm over all dims but 0 (torch weight_norm dim=0;
    for ConvTranspose1d dim 0 is Cin -- SURVEY.md a19).  Accepts the legacy
    ``weight_g/weight_v`` and the ``parametrizations.weight.original{0,1}``
    spellings and returns plain ``.weight`` entries."""
    out: Dict[str, Tensor] = {}
    for k, v in sd.items():
        if k.endswith(".weight_g") or k.endswith(".parametrizations.weight.original0"):
            continue
        if k.endswith(".weight_v") or k.endswith(".parametrizations.weight.original1"):
            if k.endswith(".weight_v"):
                base = k[: -len(".weight_v")]
                g = sd[base + ".weight_g"]
            else:
                base = k[: -len(".parametrizations.weight.original1")]
                g = sd[base + ".parametrizations.weight.original0"]
            v32 = v.float()
            nrm = v32.reshape(v32.shape[0], -1).norm(dim=1).reshape([-1] + [1] * (v32.dim() - 1))
            out[base + ".weight"] = v32 * (g.float() / nrm)
        else:
            out[k] = v.float()
    return out


def sequence_mask(lengths: Tensor, T: int) -> Tensor:
    """commons.py:89-93: arange(T) < length."""
    return (torch.arange(T, device=lengths.device)[None, :] < lengths[:, None])


# ---------------------------------------------------------------------------
# TextEncoder  (encoders.py:111-126, :61-73; attentions.py:63-113, :195-203)
# ---------------------------------------------------------------------------
def _layer_norm_c(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> Tensor:
    # normalization.py:13-16: LN over the channel dim of (B, C, T)
    mu = x.mean(dim=1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * gamma[None, :, None] + beta[None, :, None]


def _rel_attention(W, p: str, x: Tensor, mask_bt: Tensor, n_heads: int, window: int) -> Tensor:
    """Windowed relative-position self attention, attentions.py:63-113.

    scores[i,j] = q_i.k_j/sqrt(d) + [|j-i|<=w] q_i.Ek[j-i+w]/sqrt(d)
    masked_fill(mask_i*mask_j == 0, -1e4); softmax_j
    out_i = sum_j p_ij v_j + sum_{|j-i|<=w} p_ij Ev[j-i+w]
    """
    B, C, T = x.shape
    d = C // n_heads
    q = F.conv1d(x, W[p + "conv_q.weight"], W[p + "conv_q.bias"])
    k = F.conv1d(x, W[p + "conv_k.weight"], W[p + "conv_k.bias"])
    v = F.conv1d(x, W[p + "conv_v.weight"], W[p + "conv_v.bias"])
    q = q.view(B, n_heads, d, T).transpose(2, 3) / math.sqrt(d)   # (B,h,T,d)
    k = k.view(B, n_heads, d, T).transpose(2, 3)
    v = v.view(B, n_heads, d, T).transpose(2, 3)
    Ek = W[p + "emb_rel_k"][0]      # (2w+1, d), shared across heads
    Ev = W[p + "emb_rel_v"][0]
    scores = q @ k.transpose(-1, -2)                             # (B,h,T,T)
    rel = q @ Ek.t()                                             # (B,h,T,2w+1)
    for r in range(2 * window + 1):
        off = r - window                                         # j - i
        n = T - abs(off)
        if n <= 0:
            continue
        i0 = max(0, -off)
        scores.diagonal(offset=off, dim1=-2, dim2=-1).add_(rel[:, :, i0:i0 + n, r])
    pair = (mask_bt[:, None, :, None] & mask_bt[:, None, None, :])
    scores = scores.masked_fill(~pair, -1e4)
    pa = torch.softmax(scores, dim=-1)
    out = pa @ v                                                 # (B,h,T,d)
    relw = pa.new_zeros(B, n_heads, T, 2 * window + 1)
    for r in range(2 * window + 1):
        off = r - window
        n = T - abs(off)
        if n <= 0:
            continue
        i0 = max(0, -off)
        relw[:, :, i0:i0 + n, r] = pa.diagonal(offset=off, dim1=-2, dim2=-1)
    out = out + relw @ Ev
    out = out.transpose(2, 3).reshape(B, C, T)
    return F.conv1d(out, W[p + "conv_o.weight"], W[p + "conv_o.bias"])


def _ffn(W, p: str, x: Tensor, m: Tensor, ksz: int) -> Tensor:
    # attentions.py:195-203 with _same_padding (:217-221), activation relu
    pl, pr = (ksz - 1) // 2, ksz // 2
    h = F.conv1d(F.pad(x * m, (pl, pr)), W[p + "conv_1.weight"], W[p + "conv_1.bias"])
    h = torch.relu(h)
    h = F.conv1d(F.pad(h * m, (pl, pr)), W[p + "conv_2.weight"], W[p + "conv_2.bias"])
    return h * m


def text_encoder(W, cfg, phone: Tensor, pitch: Optional[Tensor], lengths: Tensor, taps=None):
    """encoders.py:111-126.  Returns m (B,C,T), logs (B,C,T), mask (B,1,T) float."""
    H = cfg.hidden_channels
    x = F.linear(phone, W["enc_p.emb_phone.weight"], W["enc_p.emb_phone.bias"])
    if pitch is not None:
        x = x + W["enc_p.emb_pitch.weight"][pitch]
    x = F.leaky_relu(x * math.sqrt(H), 0.1)
    x = x.transpose(1, 2)                                        # (B,H,T)
    T = x.shape[2]
    mask_bt = sequence_mask(lengths, T)
    m = mask_bt[:, None, :].to(x.dtype)
    x = x * m                                                    # encoders.py:122 and :63
    if taps is not None:
        taps["enc.x0"] = x.clone()
    for i in range(cfg.n_layers):
        pa = f"enc_p.encoder.attn_layers.{i}."
        y = _rel_attention(W, pa, x, mask_bt, cfg.n_heads, cfg.attn_window)
        x = _layer_norm_c(x + y, W[f"enc_p.encoder.norm_layers_1.{i}.gamma"],
                          W[f"enc_p.encoder.norm_layers_1.{i}.beta"])
        y = _ffn(W, f"enc_p.encoder.ffn_layers.{i}.", x, m, cfg.kernel_size)
        x = _layer_norm_c(x + y, W[f"enc_p.encoder.norm_layers_2.{i}.gamma"],
                          W[f"enc_p.encoder.norm_layers_2.{i}.beta"])
        if taps is not None:
            taps[f"enc.layer{i}"] = x.clone()
    x = x * m
    stats = F.conv1d(x, W["enc_p.proj.weight"], W["enc_p.proj.bias"]) * m
    mean, logs = torch.split(stats, cfg.inter_channels, dim=1)
    return mean, logs, m


# ---------------------------------------------------------------------------
# flow, reverse direction  (residuals.py:144-157, :210-229; modules.py:58-84)
# ---------------------------------------------------------------------------
def _wavenet(W, p: str, cfg, h: Tensor, m: Tensor, g: Tensor) -> Tensor:
    H = cfg.hidden_channels
    nl, kw = cfg.flow_wn_layers, cfg.flow_wn_kernel
    gc = F.conv1d(g, W[p + "cond_layer.weight"], W[p + "cond_layer.bias"])   # (B, 2H*nl, 1)
    skip = torch.zeros_like(h)
    for l in range(nl):
        a = F.conv1d(h, W[p + f"in_layers.{l}.weight"], W[p + f"in_layers.{l}.bias"],
                     padding=(kw - 1) // 2)                                    # dilation_rate 1
        a = a + gc[:, l * 2 * H:(l + 1) * 2 * H, :]
        acts = torch.tanh(a[:, :H]) * torch.sigmoid(a[:, H:])                  # commons.py:79-86
        rs = F.conv1d(acts, W[p + f"res_skip_layers.{l}.weight"], W[p + f"res_skip_layers.{l}.bias"])
        if l < nl - 1:
            h = (h + rs[:, :H]) * m
            skip = skip + rs[:, H:]
        else:
            skip = skip + rs
    return skip * m


def flow_reverse(W, cfg, z_p: Tensor, m: Tensor, g: Tensor, taps=None) -> Tensor:
    half = cfg.inter_channels // 2
    x = z_p
    for f in reversed(range(cfg.flow_n_flows)):
        x = torch.flip(x, [1])                                   # Flip precedes each RCL in reverse
        p = f"flow.flows.{2 * f}."
        x0, x1 = x[:, :half], x[:, half:]
        h = F.conv1d(x0, W[p + "pre.weight"], W[p + "pre.bias"]) * m
        h = _wavenet(W, p + "enc.", cfg, h, m, g)
        mean = F.conv1d(h, W[p + "post.weight"], W[p + "post.bias"]) * m       # mean_only: logs == 0
        x1 = (x1 - mean) * m
        x = torch.cat([x0, x1], 1)
        if taps is not None:
            taps[f"flow.{f}"] = x.clone()
    return x


# ---------------------------------------------------------------------------
# harmonic source  (generators.py:117-156, nsf.py:36-40)
# ---------------------------------------------------------------------------
def sine_source(W, cfg, f0: Tensor, eps_src: Optional[Tensor], exact: bool = False):
    """Returns (har_source (B, L, 1), sine_deterministic (B, L, 1)).

    exact=False restates the reference's op sequence (fp32 tensors, torch's
    cumsum / interpolate), so it tracks the reference to rounding.
    exact=True evaluates the closed form in float64:
        0.1 * sin(2*pi*frac(sum_{j<=n} fl32(f0[j//upp]/sr % 1))) * uv
     and is what the <=1e-5 sine check compares with.
    """
    upp, sr = cfg.upp, float(cfg.sr)
    B, T = f0.shape
    L = T * upp
    rad = (f0 / sr) % 1                                           # (B,T) fp32
    uv = (f0 > 0).to(f0.dtype)
    uv_up = uv.repeat_interleave(upp, dim=1)
    if exact:
        rad_up = rad.double().repeat_interleave(upp, dim=1)
        phase = torch.cumsum(rad_up, dim=1)
        sine = (torch.sin(2 * math.pi * (phase - torch.floor(phase))) * 0.1).float()
    else:
        over = torch.cumsum(rad, dim=1) * upp                     # generators.py:130-131
        over = F.interpolate(over[:, None, :], scale_factor=float(upp), mode="linear",
                             align_corners=True)[:, 0]            # (B,L)
        rad_up = rad.repeat_interleave(upp, dim=1)                # nearest == repeat (Appendix B)
        over = over % 1
        wrap = (over[:, 1:] - over[:, :-1]) < 0
        shift = torch.zeros_like(rad_up)
        shift[:, 1:] = wrap * -1.0
        sine = torch.sin(torch.cumsum(rad_up + shift, dim=1) * 2 * torch.pi) * 0.1
    sine = sine * uv_up
    noise_amp = uv_up * 0.003 + (1 - uv_up) * 0.1 / 3
    eps = torch.zeros(B, L) if eps_src is None else eps_src.reshape(B, L)
    wav = sine + noise_amp * eps
    w = W["dec.m_source.l_linear.weight"].reshape(())
    b = W["dec.m_source.l_linear.bias"].reshape(())
    return torch.tanh(wav * w + b)[:, :, None], sine[:, :, None]


# ---------------------------------------------------------------------------
# GeneratorNSF  (nsf.py:120-144, residuals.py:45-53)
# ---------------------------------------------------------------------------
def _resblock(W, p: str, x: Tensor, ksz: int, dils) -> Tensor:
    for d, dil in enumerate(dils):
        xt = F.leaky_relu(x, 0.1)
        xt = F.conv1d(xt, W[p + f"convs1.{d}.weight"], W[p + f"convs1.{d}.bias"],
                      dilation=dil, padding=(ksz * dil - dil) // 2)
        xt = F.leaky_relu(xt, 0.1)
        xt = F.conv1d(xt, W[p + f"convs2.{d}.weight"], W[p + f"convs2.{d}.bias"],
                      padding=(ksz - 1) // 2)
        x = xt + x
    return x


def generator(W, cfg, z: Tensor, har_source: Tensor, g: Tensor, taps=None) -> Tensor:
    src = har_source.transpose(1, 2)                              # (B,1,L)
    x = F.conv1d(z, W["dec.conv_pre.weight"], W["dec.conv_pre.bias"], padding=3)
    x = x + F.conv1d(g, W["dec.cond.weight"], W["dec.cond.bias"])
    if taps is not None:
        taps["dec.conv_pre"] = x.clone()
    nk = len(cfg.resblock_kernel_sizes)
    strides = cfg.noise_strides()
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.leaky_relu(x, 0.1)
        x = F.conv_transpose1d(x, W[f"dec.ups.{i}.weight"], W[f"dec.ups.{i}.bias"],
                               stride=u, padding=(k - u) // 2)
        s = strides[i]
        x = x + F.conv1d(src, W[f"dec.noise_convs.{i}.weight"], W[f"dec.noise_convs.{i}.bias"],
                         stride=s, padding=(s // 2 if s > 1 else 0))
        if taps is not None:
            taps[f"dec.ups{i}"] = x.clone()
        acc = None
        for j in range(nk):
            y = _resblock(W, f"dec.resblocks.{i * nk + j}.", x,
                          cfg.resblock_kernel_sizes[j], cfg.resblock_dilation_sizes[j])
            acc = y if acc is None else acc + y
        x = acc / nk
        if taps is not None:
            taps[f"dec.stage{i}"] = x.clone()
    x = F.leaky_relu(x)                                           # default slope 0.01 (nsf.py:142)
    x = torch.tanh(F.conv1d(x, W["dec.conv_post.weight"], None, padding=3))
    return x


# ---------------------------------------------------------------------------
# Synthesizer.infer
# ---------------------------------------------------------------------------
@torch.no_grad()
def infer(sd: Dict[str, Tensor], cfg, phone: Tensor, lengths: Tensor, pitch: Tensor,
          nsff0: Tensor, sid: Tensor, eps_zp: Optional[Tensor] = None,
          eps_src: Optional[Tensor] = None, taps: Optional[dict] = None,
          folded: bool = False):
    """synthesizers.py:162-188 with the two normal draws supplied by the caller
    (None == zeros).  Returns (o, x_mask, (z, z_p, m_p, logs_p)) in the
    reference's layouts."""
    W = sd if folded else fold_weight_norm(sd)
    g = W["emb_g.weight"][sid][:, :, None]                        # (B,gin,1)
    m_p, logs_p, m = text_encoder(W, cfg, phone.float(), pitch, lengths, taps)
    eps = torch.zeros_like(m_p) if eps_zp is None else eps_zp
    z_p = (m_p + torch.exp(logs_p) * eps * 0.66666) * m
    z = flow_reverse(W, cfg, z_p, m, g, taps)
    src, sine = sine_source(W, cfg, nsff0.float(), eps_src)
    if taps is not None:
        taps["source"] = src.clone()
        taps["sine"] = sine.clone()
    o = generator(W, cfg, z * m, src, g, taps)
    return o, m, (z, z_p, m_p, logs_p)
