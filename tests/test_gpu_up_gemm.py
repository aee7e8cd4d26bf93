"""rvo_up_linear (hand-written tcgen05 GEMM of the update operator's Linear layers) against a plain
PyTorch fp32 reference of the same op: y = act(x @ w.T + b) with fp16 operands (ramp/net.py:36-67 under
autocast).  Tolerance: fp32 accumulation over K = 384 then one fp16 rounding -> 2^-10 relative to the
row scale plus 2e-3 absolute slack for cancellation."""
import ctypes

import pytest
import torch

from rampvo_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(M, N, relu, ldx_pad=0, bias=True, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    K = 384
    xbuf = torch.randn(max(M, 1), K + ldx_pad, device="cuda", generator=g).half()
    x = xbuf[:M, :K]
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    b = torch.randn(N, device="cuda", generator=g).half() if bias else None
    y = torch.full((max(M, 1), N), float("nan"), dtype=torch.float16, device="cuda")
    L = _lib.lib()
    _lib.check(L.rvo_up_linear(_lib.ptr(x), x.stride(0) if M else K, _lib.ptr(w), _lib.ptr(b), M, K, N, int(relu),
                               _lib.ptr(y), y.stride(0), _lib.stream_ptr()), "rvo_up_linear")
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t()
    if bias:
        ref = ref + b.float()
    if relu:
        ref = ref.clamp_min(0)
    return y[:M], ref


@pytest.mark.parametrize("M", [1, 127, 128, 129, 5000, 45312])
@pytest.mark.parametrize("N,relu", [(384, False), (384, True), (768, False)])
def test_up_linear_matches_fp32_reference(M, N, relu):
    y, ref = _run(M, N, relu)
    assert torch.isfinite(y.float()).all()
    err = (y.float() - ref).abs()
    tol = 2.0 ** -10 * ref.abs() + 2e-3
    assert (err <= tol).all(), float((err - tol).max())


def test_up_linear_padded_rows_no_bias_and_empty():
    y, ref = _run(777, 384, False, ldx_pad=8, bias=False)
    assert ((y.float() - ref).abs() <= 2.0 ** -10 * ref.abs() + 2e-3).all()
    _run(0, 384, False)           # M = 0 is a no-op


def test_up_linear_rejects_unsupported_shapes():
    L = _lib.lib()
    x = torch.zeros(8, 128, dtype=torch.float16, device="cuda")
    w = torch.zeros(384, 128, dtype=torch.float16, device="cuda")
    y = torch.zeros(8, 384, dtype=torch.float16, device="cuda")
    rc = L.rvo_up_linear(_lib.ptr(x), 128, _lib.ptr(w), None, 8, 128, 384, 0, _lib.ptr(y), 384, _lib.stream_ptr())
    assert rc != 0 and b"K must be 384" in L.rvo_last_error()


@pytest.mark.parametrize("M", [129, 5000])
def test_up_linear_gather_equals_linear_of_gathered_rows(M):
    """rvo_up_linear_gather: input row r = x[gather[r]], a zero row for gather[r] < 0 (mask_ix * net[:, ix],
    ramp/net.py:78-82) — bit-identical to materialising the gathered matrix first"""
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(M, 384, device="cuda", generator=g).half()
    w = (torch.randn(384, 384, device="cuda", generator=g) / 384 ** 0.5).half()
    b = torch.randn(384, device="cuda", generator=g).half()
    idx = torch.randint(-1, M, (M,), device="cuda", dtype=torch.int64, generator=g)
    xg = torch.where((idx >= 0)[:, None], x[idx.clamp(min=0)], torch.zeros(1, dtype=torch.float16, device="cuda"))
    L = _lib.lib()
    ya = torch.empty(M, 384, dtype=torch.float16, device="cuda")
    yb = torch.empty(M, 384, dtype=torch.float16, device="cuda")
    _lib.check(L.rvo_up_linear_gather(_lib.ptr(x), 384, _lib.ptr(idx), _lib.ptr(w), _lib.ptr(b), M, 384, 384, 1,
                                      _lib.ptr(ya), 384, _lib.stream_ptr()), "rvo_up_linear_gather")
    _lib.check(L.rvo_up_linear(_lib.ptr(xg), 384, _lib.ptr(w), _lib.ptr(b), M, 384, 384, 1, _lib.ptr(yb), 384,
                               _lib.stream_ptr()), "rvo_up_linear")
    torch.cuda.synchronize()
    assert torch.equal(ya, yb)
    ref = (xg.float() @ w.float().t() + b.float()).clamp_min(0)
    assert ((ya.float() - ref).abs() <= 2.0 ** -10 * ref.abs() + 2e-3).all()
