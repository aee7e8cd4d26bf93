"""GPU: the tcgen05 implicit-GEMM convolution (rvo_conv2d_nhwc, csrc/conv_tc.cu) against a plain PyTorch fp32
reference of the same op on the same fp16-rounded operands — every layer shape of the RAMP encoder CNNs
(ramp/extractor.py:60-130,272-311) plus ragged sizes (partial tiles, borders narrower than the kernel)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from rampvo_b200.extractor import CL, _conv_tc

pytestmark = pytest.mark.gpu

SHAPES = [
    # C0, C1, Cout, ks, stride, pad, H, W
    (16, 0, 32, 7, 2, 3, 480, 640),      # conv1
    (32, 0, 32, 3, 1, 1, 240, 320),      # layer1
    (32, 32, 64, 3, 2, 1, 240, 320),     # layer3.0.conv1 on cat(x, x_down2)
    (32, 32, 64, 1, 2, 0, 240, 320),     # layer3.0.downsample
    (64, 0, 64, 3, 1, 1, 120, 160),      # layer3.*
    (64, 64, 128, 1, 1, 0, 120, 160),    # conv3 (fmap)
    (64, 64, 384, 1, 1, 0, 120, 160),    # conv3 (imap)
    (16, 0, 32, 7, 2, 3, 17, 23),        # ragged: one partial tile, image narrower than the padding ring reaches
    (8, 8, 48, 3, 1, 1, 5, 300),
    (64, 0, 64, 3, 2, 1, 31, 33),
    (32, 0, 16, 5, 1, 2, 40, 40),
]


@pytest.mark.parametrize("C0,C1,Cout,ks,stride,pad,H,W", SHAPES)
def test_conv_tc_matches_fp32_reference(C0, C1, Cout, ks, stride, pad, H, W):
    g = torch.Generator(device="cuda").manual_seed(C0 * 1000 + Cout + H)
    conv = nn.Conv2d(C0 + C1, Cout, ks, stride=stride, padding=pad).cuda()
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g, device="cuda") / (ks * ks * (C0 + C1)) ** 0.5)
        conv.bias.copy_(torch.randn(Cout, generator=g, device="cuda"))
    x = torch.randn(1, C0, H, W, generator=g, device="cuda").half().contiguous(memory_format=CL)
    x2 = torch.randn(1, C1, H, W, generator=g, device="cuda").half().contiguous(memory_format=CL) if C1 else None
    got, st = _conv_tc(conv, x, x2, stats=True)
    xin = x.float() if x2 is None else torch.cat((x.float(), x2.float()), dim=1)
    ref = F.conv2d(xin, conv.weight.half().float(), conv.bias.float(), stride=stride, padding=pad)
    assert got.shape == ref.shape and got.dtype == torch.float16 and got.is_contiguous(memory_format=CL)
    d = (got.float() - ref).abs()
    tol = 2.0 ** -10 * ref.abs() + 2e-3            # one fp16 rounding of the output + fp32 summation order
    assert bool((d <= tol).all()), "max |d| %.3e" % d.max().item()
    got2, st2 = _conv_tc(conv, x, x2, stats=True)      # statistics buffers are zeroed by the caller side, per call
    assert torch.equal(got, got2) and torch.allclose(st, st2, rtol=1e-5)
    # InstanceNorm statistics of the rounded outputs
    s1 = got.float().sum(dim=(0, 2, 3))
    s2 = (got.float() ** 2).sum(dim=(0, 2, 3))
    n = got.shape[2] * got.shape[3]
    assert torch.allclose(st[:Cout], s1, rtol=1e-4, atol=1e-3 * n ** 0.5)
    assert torch.allclose(st[Cout:], s2, rtol=1e-4, atol=1e-3 * n ** 0.5)


def test_conv_tc_scale_and_weight_cache_invalidation():
    conv = nn.Conv2d(32, 32, 3, padding=1).cuda()
    x = torch.randn(1, 32, 24, 40, device="cuda").half().contiguous(memory_format=CL)
    a, _ = _conv_tc(conv, x)
    b, _ = _conv_tc(conv, x, scale=0.25)
    assert (a.float() * 0.25 - b.float()).abs().max().item() < 2e-3
    with torch.no_grad():
        conv.weight.mul_(2.0)                    # in-place update bumps the version counter: cache must rebuild
        conv.bias.mul_(2.0)
    c, _ = _conv_tc(conv, x)
    assert (a.float() * 2 - c.float()).abs().max().item() < 1e-2
