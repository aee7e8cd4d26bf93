"""GPU: the training-time gradients of altcorr (rvo_corr_backward / rvo_patchify_backward) against
the reference's own compiled backward ops (cuda_corr.backward / patchify_backward from oracle/_ref) and through
autograd against the reference's CorrLayer / PatchLayer (ramp/altcorr/correlation.py:4-68)."""
import numpy as np
import pytest
import torch

from oracle import build_ref
from oracle import ref_gpu_vo as R
from rampvo_b200 import altcorr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_corr():
    m = build_ref.load_ref("cuda_corr_ref")
    if m is None:
        pytest.skip("oracle/_ref/cuda_corr_ref.so not built")
    return m


def _problem(C, R, E=300, Np=40, Nf=6, H=30, W=40, P=3, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    f1 = torch.randn(1, Np, C, P, P, generator=g, device="cuda") / C ** 0.5
    f2 = torch.randn(1, Nf, C, H, W, generator=g, device="cuda") / C ** 0.5
    ii = torch.randint(0, Np, (E,), generator=g, device="cuda")
    jj = torch.randint(0, Nf, (E,), generator=g, device="cuda")
    ctr = torch.stack([torch.rand(E, generator=g, device="cuda") * (W + 8) - 4,
                       torch.rand(E, generator=g, device="cuda") * (H + 8) - 4], 1)        # some windows leave the map
    off = torch.arange(P, device="cuda", dtype=torch.float32) - P // 2
    coords = torch.zeros(1, E, 2, P, P, device="cuda")
    coords[0, :, 0] = ctr[:, 0, None, None] + off[None, None, :] * 1.03
    coords[0, :, 1] = ctr[:, 1, None, None] + off[None, :, None] * 0.97
    d = 2 * R + 1
    grad = torch.randn(1, E, d, d, P, P, generator=g, device="cuda")
    return f1, f2, coords, ii, jj, grad


@pytest.mark.parametrize("C,R", [(128, 3), (128, 1), (48, 3), (8, 0)])
def test_corr_backward_matches_reference_op(ref_corr, C, R):
    f1, f2, coords, ii, jj, grad = _problem(C, R, seed=C + R)
    r1, r2 = ref_corr.backward(f1, f2, coords, ii, jj, grad, R)
    g1, g2 = altcorr.corr_backward(f1, f2, coords, ii, jj, grad, R)
    for ours, ref in ((g1, r1), (g2, r2)):
        assert ours.shape == ref.shape and ours.dtype == ref.dtype
        err = (ours - ref).abs().max().item() / ref.abs().max().item()
        assert err < 2e-5, err                     # fp32 atomics on both sides: accumulation-order noise


def test_corr_backward_half_inputs(ref_corr):
    f1, f2, coords, ii, jj, grad = _problem(128, 3, seed=5)
    r1, r2 = ref_corr.backward(f1.half().float(), f2.half().float(), coords, ii, jj, grad, 3)
    g1, g2 = altcorr.corr_backward(f1.half(), f2.half(), coords, ii, jj, grad, 3)
    assert g1.dtype == torch.float16 and g2.dtype == torch.float16
    assert (g1.float() - r1).abs().max().item() < 2e-3 * r1.abs().max().item()
    assert (g2.float() - r2).abs().max().item() < 2e-3 * r2.abs().max().item()


@pytest.mark.parametrize("dropout", [1, 0.5])
def test_corr_autograd_matches_reference_layer(dropout):
    if not R.available():
        pytest.skip("oracle/_ref not built")
    ns = R.load(with_vo=False)
    f1, f2, coords, ii, jj, grad = _problem(128, 3, seed=9)
    outs = []
    for fn in (ns.altcorr.corr, altcorr.corr):
        a, b = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
        torch.manual_seed(3)                       # the dropout mask is drawn in backward (correlation.py:20-25)
        y = fn(a, b, coords, ii, jj, 3, dropout)
        y.backward(grad)
        outs.append((y.detach(), a.grad, b.grad))
    (y0, a0, b0), (y1, a1, b1) = outs
    assert (y0 - y1).abs().max().item() < 2e-6 * max(1.0, y0.abs().max().item())
    assert (a0 - a1).abs().max().item() < 2e-5 * a0.abs().max().item()
    assert (b0 - b1).abs().max().item() < 2e-5 * b0.abs().max().item()


@pytest.mark.parametrize("C,R,dtype", [(128, 1, torch.float32), (3, 1, torch.float32), (384, 0, torch.float32),
                                      (128, 1, torch.float16)])
def test_patchify_backward_matches_reference_op(ref_corr, C, R, dtype):
    g = torch.Generator(device="cuda").manual_seed(C)
    B, H, W, M = 2, 30, 40, 96
    net = torch.randn(B, C, H, W, generator=g, device="cuda").to(dtype)
    coords = torch.stack([torch.rand(B, M, generator=g, device="cuda") * (W + 4) - 2,
                          torch.rand(B, M, generator=g, device="cuda") * (H + 4) - 2], -1)
    D = 2 * R + 2
    grad = torch.randn(B, M, C, D, D, generator=g, device="cuda").to(dtype)
    ref, = ref_corr.patchify_backward(net, coords, grad, R)
    got, = altcorr.patchify_backward(net.shape, net.dtype, coords, grad, R)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    tol = 2e-5 if dtype == torch.float32 else 2e-2      # the reference accumulates fp16 gradients in fp16 atomics
    assert (got.float() - ref.float()).abs().max().item() < tol * max(1.0, ref.float().abs().max().item())


def test_patchify_autograd_matches_reference_layer():
    if not R.available():
        pytest.skip("oracle/_ref not built")
    ns = R.load(with_vo=False)
    g = torch.Generator(device="cuda").manual_seed(11)
    net = torch.randn(1, 128, 30, 40, generator=g, device="cuda")
    coords = torch.stack([torch.rand(1, 96, generator=g, device="cuda") * 38 + 1,
                          torch.rand(1, 96, generator=g, device="cuda") * 28 + 1], -1)
    w = torch.randn(1, 96, 128, 3, 3, generator=g, device="cuda")
    res = []
    for fn in (ns.altcorr.patchify, altcorr.patchify):
        a = net.clone().requires_grad_(True)
        y = fn(a, coords, 1)
        (y * w).sum().backward()
        res.append((y.detach(), a.grad))
    assert torch.equal(res[0][0], res[1][0])
    assert (res[0][1] - res[1][1]).abs().max().item() < 2e-5 * res[0][1].abs().max().item()
