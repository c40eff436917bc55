"""The C-ABI shared library: loads without a GPU, exports every symbol the header declares,
struct layout matches the ctypes mirror, and fails loudly (no fallback) when no device exists."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "polgen_rvc.h")


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, os.path.join(ROOT, "polgen-rvc_b200"))
    import polgen_rvc_b200  # noqa: F401
    from polgen_rvc_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("pg_build", os.path.join(ROOT, "polgen-rvc_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return _lib


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    names = declared_symbols()
    assert "pg_infer" in names and "pg_create" in names and len(names) >= 15
    handle = C.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in polgen_rvc.h but not exported"
    assert set(names) == set(lib.SYMBOLS), "ctypes table and header disagree"


def test_abi_version(lib):
    assert lib.load().pg_abi_version() == lib.PG_ABI_VERSION == 2


def test_config_struct_layout_matches_header(lib):
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "polgen_rvc.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(pg_config), offsetof(pg_config, upsample_rates),
         offsetof(pg_config, resblock_dilations), offsetof(pg_config, flow_n_flows), offsetof(pg_config, flags));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    size, o_rates, o_dil, o_flows, o_flags = (int(v) for v in out)
    S = lib.PgConfig
    assert C.sizeof(S) == size
    assert S.upsample_rates.offset == o_rates
    assert S.resblock_dilations.offset == o_dil
    assert S.flow_n_flows.offset == o_flows
    assert S.flags.offset == o_flags


def test_segment_struct_layout_matches_header(lib):
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "polgen_rvc.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(pg_segment), offsetof(pg_segment, trim), offsetof(pg_segment, sid),
         offsetof(pg_segment, phone), offsetof(pg_segment, wave), offsetof(pg_segment, aux));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    S = lib.PgSegment
    assert [C.sizeof(S), S.trim.offset, S.sid.offset, S.phone.offset, S.wave.offset, S.aux.offset] == out


def test_argument_errors_without_device(lib):
    L = lib.load()
    assert L.pg_create(None, 0, None) == -1
    assert b"null" in L.pg_last_error()
    assert L.pg_finalize(None) == -1
    assert L.pg_workspace_bytes(None, 1, 1) == 0
    assert L.pg_destroy(None) == 0
    assert L.pg_infer_segments(None, None, 0, None, 0) == -1
    assert L.pg_postprocess(None, None, None, 0, None, 0, 16000, 48000, 1.0, None, None) == -1
    assert L.pg_coarse_pitch(None, None, 0, None, 0, 50.0, 1100.0, None, None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_error_not_fallback(lib):
    import polgen_rvc_b200 as pg
    with pytest.raises(lib.PgError):
        pg.Engine(pg.CONFIGS["v2-40k"], {}, 0)
    net = pg.Synthesizer(*pg.CONFIGS["v2-40k"].ctor_args(), use_f0=1, input_dim=768, is_half=False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.infer(torch.zeros(1, 4, 768), torch.tensor([4]), torch.ones(1, 4).long(), torch.ones(1, 4),
                  torch.zeros(1).long())


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package may import it."""
    pkg = os.path.join(ROOT, "polgen-rvc_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle|rvc_oracle|oracle/", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f"{f} references the oracle"


def test_flag_constants_match_header(lib):
    """every PG_FLAG_* of the header has the same value in the ctypes mirror (the twins the GPU tests select), and no
    two flags share a bit"""
    src = open(HEADER).read()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(PG_FLAG_[A-Z0-9_]+)\s+(\d+)", src)}
    assert len(flags) >= 12
    for name, value in flags.items():
        assert getattr(lib, name) == value, name
        assert value & (value - 1) == 0, f"{name} is not a single bit"
    assert len(set(flags.values())) == len(flags)
