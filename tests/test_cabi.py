"""CPU: the C-ABI library builds, loads and exports every symbol include/rampvo_b200.h declares."""
import ctypes
import os
import re

from rampvo_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rampvo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rvo_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    path = build.build_library()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "missing export %s" % s


def test_binding_covers_header():
    assert sorted(_lib.exported_symbols()) == header_symbols()


def test_version_and_error_string(lib):
    assert lib.rvo_abi_version() == 1
    assert lib.rvo_plan_bytes(0) > 0
    assert lib.rvo_plan_bytes(45312) > 45312 * 4
    assert lib.rvo_ba_ws_bytes(45312, 196608, 10) > lib.rvo_plan_bytes(45312)
    assert lib.rvo_plan_bytes(-1) == -1


def test_argument_errors_do_not_abort(lib):
    # bad arguments return an error code and a message (the reference exit(1)s, block_e.cu:20-26)
    rc = lib.rvo_transform(None, None, None, None, None, None, 5, 3, 0, None, None, None, None, None)
    assert rc == 1
    assert b"null" in lib.rvo_last_error()
    rc = lib.rvo_ba_forward(None, None, None, None, None, None, None, None, None, -1, 1, 1, 3, 96,
                            0, 1, 2, 0, None, 0, None)
    assert rc == 1


def test_no_cpu_fallback():
    import pytest
    import torch
    from rampvo_b200 import altcorr
    with pytest.raises(RuntimeError):
        altcorr.corr(torch.zeros(1, 1, 8, 3, 3), torch.zeros(1, 1, 8, 4, 4),
                     torch.zeros(1, 1, 2, 3, 3), torch.zeros(1, dtype=torch.long),
                     torch.zeros(1, dtype=torch.long), 1)


def test_up_linear_argument_errors(lib):
    # shape / pointer validation happens before any CUDA call (no GPU needed)
    rc = lib.rvo_up_linear(None, 384, None, None, 8, 128, 384, 0, None, 384, None)
    assert rc == 1 and b"K must be 384" in lib.rvo_last_error()
    rc = lib.rvo_up_linear(None, 384, None, None, 8, 384, 100, 0, None, 384, None)
    assert rc == 1
    rc = lib.rvo_up_linear(None, 384, None, None, 8, 384, 384, 0, None, 384, None)
    assert rc == 1 and b"null" in lib.rvo_last_error()
    assert lib.rvo_up_linear(None, 384, None, None, 0, 384, 384, 0, None, 384, None) == 0     # M = 0: no-op


def test_corr_tiles_workspace_size(lib):
    import ctypes
    arr = (_lib.FMap * 2)(_lib.FMap(0, _lib.RVO_F16, 32, 128, 120, 160, 120 * 160 * 128, 1, 160 * 128, 128),
                          _lib.FMap(0, _lib.RVO_F16, 32, 128, 30, 40, 30 * 40 * 128, 1, 40 * 128, 128))
    n0 = lib.rvo_corr_tiles_ws_bytes(arr, 2, 0)
    n1 = lib.rvo_corr_tiles_ws_bytes(arr, 2, 45312)
    # a 32-byte record per row + one block header per 128 rows and per tile + the (tile, window-origin row) counters
    assert 0 < n0 < n1 and 45312 * 18 * 32 <= n1 <= 45312 * 18 * 40
    assert lib.rvo_corr_tiles_ws_bytes(arr, 2, -1) == -1
