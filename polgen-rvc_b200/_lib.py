"""ctypes binding of libpolgen_rvc.so (the C ABI in include/polgen_rvc.h).

There is no fallback: if the shared library has not been built
(``python polgen-rvc_b200/build.py`` or ``__graft_entry__.build()``) importing
this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PG_LIB_PATH") or os.path.join(_HERE, "libpolgen_rvc.so")   # PG_LIB_PATH: A/B builds

PG_MAX_UPS = 8
PG_MAX_RESBLOCK_KERNELS = 4
PG_MAX_DILATIONS = 4
PG_FLAG_FORCE_SIMT = 1
PG_FLAG_KEEP_TAPS = 2
PG_FLAG_PROFILE = 4
PG_FLAG_F16_LATENTS = 8
PG_FLAG_LEGACY_DECODER = 16
PG_FLAG_NO_PAIR_FUSION = 32
PG_FLAG_NO_GRAPHS = 64
PG_FLAG_NO_PLANES_SWAP = 128
PG_FLAG_NO_PAD = 256
PG_FLAG_F32_STREAM = 512
PG_FLAG_NO_NOISE_FUSION = 1024
PG_FLAG_LEGACY_ATTENTION = 2048
PG_FLAG_NO_POST_FUSION = 4096
PG_F32 = 0
PG_F64 = 3
PG_ABI_VERSION = 2
PROFILE_COLS = 9     # pg_profile_table row: class, Cin, N, K, dil, MT, launches, ms, flops


class PgConfig(C.Structure):
    _fields_ = [
        ("input_dim", C.c_int32), ("inter_channels", C.c_int32), ("hidden_channels", C.c_int32),
        ("filter_channels", C.c_int32), ("n_heads", C.c_int32), ("n_layers", C.c_int32),
        ("kernel_size", C.c_int32), ("attn_window", C.c_int32), ("gin_channels", C.c_int32),
        ("spk_embed_dim", C.c_int32), ("sr", C.c_int32), ("upsample_initial_channel", C.c_int32),
        ("n_ups", C.c_int32), ("upsample_rates", C.c_int32 * PG_MAX_UPS),
        ("upsample_kernel_sizes", C.c_int32 * PG_MAX_UPS), ("n_resblock_kernels", C.c_int32),
        ("resblock_kernel_sizes", C.c_int32 * PG_MAX_RESBLOCK_KERNELS), ("n_dilations", C.c_int32),
        ("resblock_dilations", (C.c_int32 * PG_MAX_DILATIONS) * PG_MAX_RESBLOCK_KERNELS),
        ("flow_n_flows", C.c_int32), ("flow_wn_layers", C.c_int32), ("flow_wn_kernel", C.c_int32),
        ("flags", C.c_int32),
    ]


class PgSegment(C.Structure):
    """pg_segment (include/polgen_rvc.h): one stand-alone segment of a pg_infer_segments call."""
    _fields_ = [
        ("T", C.c_int32), ("trim", C.c_int32), ("sid", C.c_int64),
        ("phone", C.c_void_p), ("pitch", C.c_void_p), ("f0", C.c_void_p),
        ("eps_zp", C.c_void_p), ("eps_src", C.c_void_p), ("wave", C.c_void_p), ("aux", C.c_void_p),
    ]


# every symbol include/polgen_rvc.h declares: name -> (restype, argtypes)
_P, _I, _F = C.c_void_p, C.c_int, C.c_float
SYMBOLS = {
    "pg_last_error": (C.c_char_p, []),
    "pg_abi_version": (_I, []),
    "pg_create": (_I, [C.POINTER(PgConfig), _I, C.POINTER(_P)]),
    "pg_load_tensor": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I, _I]),
    "pg_finalize": (_I, [_P]),
    "pg_workspace_bytes": (C.c_size_t, [_P, _I, _I]),
    "pg_infer": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, C.c_uint64, _P, _P]),
    "pg_infer_segments": (_I, [_P, _P, _I, C.POINTER(PgSegment), C.c_uint64]),
    "pg_infer_host": (_I, [_P, _I, _I, _P, _P, _P, _P, _P, C.c_uint64, _P]),
    "pg_coarse_pitch": (_I, [_P, _P, C.c_int64, _P, _I, C.c_double, C.c_double, _P, _P]),
    "pg_prepare_features": (_I, [_P, _P, C.c_int64, C.c_int64, _P, _P, _P, _F, _P, C.POINTER(C.c_int64)]),
    "pg_postprocess": (_I, [_P, _P, _P, C.c_int64, _P, C.c_int64, _I, _I, _F, _P, _P]),
    "pg_text_encoder": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "pg_flow_reverse": (_I, [_P, _P, _I, _I, _P, _P, _P, _P]),
    "pg_source": (_I, [_P, _P, _I, _I, _P, _P, C.c_uint64, _P, _P]),
    "pg_generator": (_I, [_P, _P, _I, _I, _P, _P, _P, _P]),
    "pg_debug_fetch": (C.c_int64, [_P, _P, C.c_char_p, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "pg_launch_count": (C.c_int64, [_P]),
    "pg_graph_count": (C.c_int, [_P]),
    "pg_set_graph_cache": (_I, [_P, _I]),
    "pg_set_decoder_sms": (_I, [_P, _I]),
    "pg_padded_frames": (_I, [_P, _I]),
    "pg_profile_read": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "pg_profile_table": (_I, [_P, C.POINTER(C.c_double), _I]),
    "pg_op_conv1d_f16": (_I, [_I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _F, _F, _P, _P, _I,
                              C.POINTER(_F)]),
    "pg_destroy": (_I, [_P]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python polgen-rvc_b200/build.py). polgen-rvc_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class PgError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pg_last_error()
        raise PgError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def make_config(cfg, flags: int = 0) -> PgConfig:
    """cfg: configs.SynthConfig."""
    c = PgConfig()
    c.input_dim = cfg.input_dim
    c.inter_channels = cfg.inter_channels
    c.hidden_channels = cfg.hidden_channels
    c.filter_channels = cfg.filter_channels
    c.n_heads = cfg.n_heads
    c.n_layers = cfg.n_layers
    c.kernel_size = cfg.kernel_size
    c.attn_window = cfg.attn_window
    c.gin_channels = cfg.gin_channels
    c.spk_embed_dim = cfg.spk_embed_dim
    c.sr = cfg.sr
    c.upsample_initial_channel = cfg.upsample_initial_channel
    if len(cfg.upsample_rates) > PG_MAX_UPS or len(cfg.resblock_kernel_sizes) > PG_MAX_RESBLOCK_KERNELS:
        raise ValueError("config exceeds the C ABI's fixed array sizes")
    c.n_ups = len(cfg.upsample_rates)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        c.upsample_rates[i] = u
        c.upsample_kernel_sizes[i] = k
    c.n_resblock_kernels = len(cfg.resblock_kernel_sizes)
    nd = len(cfg.resblock_dilation_sizes[0])
    if nd > PG_MAX_DILATIONS or any(len(d) != nd for d in cfg.resblock_dilation_sizes):
        raise ValueError("resblock dilation lists must have equal length <= 4")
    c.n_dilations = nd
    for j, (k, dil) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
        c.resblock_kernel_sizes[j] = k
        for d, v in enumerate(dil):
            c.resblock_dilations[j][d] = v
    c.flow_n_flows = cfg.flow_n_flows
    c.flow_wn_layers = cfg.flow_wn_layers
    c.flow_wn_kernel = cfg.flow_wn_kernel
    c.flags = flags
    return c
