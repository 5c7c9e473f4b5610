"""The C-ABI boundary without a GPU: the library loads, exports every symbol the header
declares with the arity the ctypes table binds, and validates arguments before touching
CUDA (so these calls are safe on a CPU-only box)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "fuxictr_b200.h")


def header_prototypes():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"B2_API\s+([\w\s\*]+?)\s*\b(b2_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(3).strip()
        nargs = 0 if args in ("void", "") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(2)] = nargs
    return protos


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    from fuxictr_b200 import _lib
    return _lib


def test_header_declares_what_python_binds(lib):
    protos = header_prototypes()
    assert len(protos) >= 17
    assert set(protos) == set(lib.SIGNATURES), set(protos) ^ set(lib.SIGNATURES)
    for name, nargs in protos.items():
        assert len(lib.SIGNATURES[name][1]) == nargs, name


def test_library_exports_every_symbol(lib):
    handle = ctypes.CDLL(lib.LIB_PATH)
    for name in header_prototypes():
        assert hasattr(handle, name), name
    assert lib.version().startswith("fuxictr_b200")


def test_field_struct_layout(lib):
    assert ctypes.sizeof(lib.b2_field) == 64
    assert lib.b2_field.vocab.offset == 24 and lib.b2_field.dim.offset == 48


def test_argument_validation_needs_no_gpu(lib):
    L = lib.load()
    null = ctypes.c_void_p(0)
    rc = L.b2_gemm_f32(null, 1, 1, null, 1, 1, null, 1, 4, 4, 4, null, 0, null, null, 0, null)
    assert rc == -1 and b"NULL" in L.b2_last_error()
    rc = L.b2_fm_fwd(ctypes.c_void_p(16), 4, 3, 8, 7, ctypes.c_void_p(16), null)
    assert rc == -1 and b"mode" in L.b2_last_error()
    fields = (lib.b2_field * 1)()
    rc = L.b2_embed_gather_fwd(fields, 0, 4, lib.B2_F64, lib.B2_F32, null, null, null)
    assert rc == -1 and b"nfields" in L.b2_last_error()
    rc = L.b2_embed_gather_fwd(fields, 1, 4, lib.B2_F64, lib.B2_BF16, null, null, null)
    assert rc == -1
    with pytest.raises(lib.B2Error):
        lib.call("b2_act_bwd", null, null, null, 8, 0, null)


def test_missing_library_is_a_hard_error(lib, monkeypatch):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libfuxictr_b200.so")
    with pytest.raises(ImportError):
        lib.load()


def test_metric_entry_points_validate_before_cuda(lib):
    """b2_auc / b2_sort_u32 / b2_logloss_sum reject bad sizes, NULLs, small or misaligned workspaces
    without touching a device; b2_auc_workspace_bytes is pure arithmetic."""
    L = lib.load()
    null = ctypes.c_void_p(0)
    nbytes = ctypes.c_int64(0)
    assert L.b2_auc_workspace_bytes(1000, ctypes.byref(nbytes)) == 0
    small = nbytes.value
    assert small >= 2 * 4 * 1000 + 256 * 4 and small % 256 == 0
    assert L.b2_auc_workspace_bytes(6_000_000, ctypes.byref(nbytes)) == 0 and nbytes.value > 48_000_000
    assert L.b2_auc_workspace_bytes(0, ctypes.byref(nbytes)) == -1
    assert L.b2_auc_workspace_bytes(1 << 31, ctypes.byref(nbytes)) == -1 and b"2^31" in L.b2_last_error()
    ptr = ctypes.c_void_p(4096)
    assert L.b2_auc(null, ptr, 10, ptr, 1 << 20, ptr, null) == -1 and b"NULL" in L.b2_last_error()
    assert L.b2_auc(ptr, ptr, 1000, ptr, small - 1, ptr, null) == -1 and b"workspace" in L.b2_last_error()
    assert L.b2_auc(ptr, ptr, 1000, ctypes.c_void_p(4096 + 8), small, ptr, null) == -1 and b"aligned" in L.b2_last_error()
    assert L.b2_auc(ptr, ptr, 1000, ptr, small, ctypes.c_void_p(4100), null) == -1
    assert L.b2_sort_u32(null, 0, null, 0, null) == 0                     # empty input: nothing to do
    assert L.b2_sort_u32(null, 5, ptr, small, null) == -1
    assert L.b2_logloss_sum(null, null, 0, null, null) == 0
    assert L.b2_logloss_sum(null, ptr, 4, ptr, null) == -1
    assert L.b2_set_l2_fetch_granularity(48) == -1 and b"granularity" in L.b2_last_error()
