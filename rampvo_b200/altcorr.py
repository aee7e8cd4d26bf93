"""Drop-in for ramp.altcorr (ramp/altcorr/correlation.py:51-74, cuda_corr ramp/altcorr/correlation.cpp:57-62).

Same names, argument order and tensor layouts as the reference; the work is done by
librampvo_b200.so (rvo_patchify_*, rvo_corr_*).  With gradients enabled `corr` and `patchify` are
autograd Functions like the reference's CorrLayer / PatchLayer (correlation.py:4-47), backed by
rvo_corr_backward / rvo_patchify_backward (csrc/altcorr_bwd.cu).
"""
import ctypes

import torch

from . import _lib


def _needs_grad(*ts):
    return torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in ts)


class _PatchLayer(torch.autograd.Function):
    """correlation.py:33-47: raw (2R+2)^2 gather with a scatter-add backward"""

    @staticmethod
    def forward(ctx, net, coords, radius):
        ctx.radius = radius
        ctx.save_for_backward(coords)
        ctx.net_shape, ctx.net_dtype = tuple(net.shape), net.dtype
        return patchify_forward(net, coords, radius)[0]

    @staticmethod
    def backward(ctx, grad):
        coords, = ctx.saved_tensors
        return patchify_backward(ctx.net_shape, ctx.net_dtype, coords, grad, ctx.radius)[0], None, None


class _CorrLayer(torch.autograd.Function):
    """correlation.py:4-30: dropout < 1 keeps a random subset of the edges in the backward pass only"""

    @staticmethod
    def forward(ctx, fmap1, fmap2, coords, ii, jj, radius, dropout):
        ctx.save_for_backward(fmap1, fmap2, coords, ii, jj)
        ctx.radius, ctx.dropout = radius, dropout
        with torch.no_grad():
            return corr(fmap1, fmap2, coords, ii, jj, radius)

    @staticmethod
    def backward(ctx, grad):
        fmap1, fmap2, coords, ii, jj = ctx.saved_tensors
        if ctx.dropout < 1:
            perm = torch.rand(len(ii), device=ii.device) < ctx.dropout
            coords, grad, ii, jj = coords[:, perm], grad[:, perm], ii[perm], jj[perm]
        g1, g2 = corr_backward(fmap1, fmap2, coords, ii, jj, grad, ctx.radius)
        return g1, g2, None, None, None, None, None


def patchify_backward(net_shape, net_dtype, coords, grad, radius):
    """cuda_corr.patchify_backward: grad [B,M,C,D,D] -> [net_grad [B,C,H,W]] in net's dtype"""
    _lib.require_cuda(coords, grad)
    B, C, H, W = net_shape
    M = coords.shape[1]
    coords = coords.to(torch.float32).contiguous()
    grad = grad.contiguous()
    if grad.dtype not in (torch.float16, torch.float32):
        grad = grad.float()
    out = torch.empty(B, C, H, W, dtype=torch.float32, device=grad.device)
    with torch.cuda.device(grad.device):
        for b in range(B):
            _lib.check(_lib.lib().rvo_patchify_backward(_lib.ptr(grad[b]), _lib.dtype_code(grad), _lib.ptr(coords[b]),
                                                        M, C, H, W, radius, _lib.ptr(out[b]), _lib.stream_ptr()),
                       "rvo_patchify_backward")
    return [out.to(net_dtype)]


def corr_backward(fmap1, fmap2, coords, ii, jj, grad, radius):
    """cuda_corr.backward: gradients w.r.t. fmap1 [B,Np,C,P,P] and fmap2 [B,Nf,C,H2,W2] given grad
    [B,E,2R+1,2R+1,P,P]; returned in the inputs' dtype like the reference's zeros_like buffers."""
    _lib.require_cuda(fmap1, fmap2, coords, ii, jj, grad)
    B, E = coords.shape[0], coords.shape[1]
    coords = coords.to(torch.float32).contiguous()
    grad = grad.to(torch.float32).contiguous()
    ii, jj = ii.to(torch.int64).contiguous(), jj.to(torch.int64).contiguous()
    g1 = torch.empty(fmap1.shape, dtype=torch.float32, device=fmap1.device)
    g2 = torch.empty(fmap2.shape, dtype=torch.float32, device=fmap2.device)
    with torch.cuda.device(fmap1.device):
        for b in range(B):
            v1, v2 = _views(fmap1[b], fmap2[b])
            _lib.check(_lib.lib().rvo_corr_backward(ctypes.byref(v1), ctypes.byref(v2), _lib.ptr(coords[b]),
                                                    _lib.ptr(ii), _lib.ptr(jj), _lib.ptr(grad[b]), E, radius,
                                                    _lib.ptr(g1[b]), _lib.ptr(g2[b]), _lib.stream_ptr()),
                       "rvo_corr_backward")
    return [g1.to(fmap1.dtype), g2.to(fmap2.dtype)]


def patchify_forward(net, coords, radius):
    """cuda_corr.patchify_forward: raw (2R+2)^2 gather, [B,M,C,D,D] in net's dtype, bit-exact."""
    _lib.require_cuda(net, coords)
    B, C, H, W = net.shape
    M = coords.shape[1]
    D = 2 * radius + 2
    coords = coords.to(torch.float32).contiguous()
    out = torch.empty(B, M, C, D, D, dtype=net.dtype, device=net.device)
    with torch.cuda.device(net.device):
        v = _lib.fmap_view(net)
        _lib.check(_lib.lib().rvo_patchify_forward(ctypes.byref(v), _lib.ptr(coords), M, radius,
                                                   _lib.ptr(out), _lib.stream_ptr()),
                   "rvo_patchify_forward")
    return [out]


def patchify(net, coords, radius, mode='bilinear', out=None):
    """ramp.altcorr.patchify (correlation.py:51-68).

    net [B,C,H,W] (any strides, fp16/fp32), coords [B,M,2] (x,y).  'bilinear' returns
    [B,M,C,2R+1,2R+1] in fp32 (the reference's fp32 offsets promote the product to fp32,
    correlation.py:57-66); any other mode returns the raw [B,M,C,2R+2,2R+2] gather.
    `out` (optional, extension): a preallocated fp16/fp32 tensor view of shape [B,M,C,d,d] with
    arbitrary strides — lets the caller write straight into a channels-last ring buffer.
    """
    if _needs_grad(net, coords):
        # training: PatchLayer + the 4-corner blend in torch ops, exactly like correlation.py:51-68 (the blend
        # weights stay differentiable w.r.t. coords through autograd)
        patches = _PatchLayer.apply(net, coords, radius)
        if mode != 'bilinear':
            return patches
        offset = (coords - coords.floor()).to(net.device)
        dx, dy = offset[:, :, None, None, None].unbind(dim=-1)
        d = 2 * radius + 1
        res = ((1 - dy) * (1 - dx) * patches[..., :d, :d] + (1 - dy) * dx * patches[..., :d, 1:] +
               dy * (1 - dx) * patches[..., 1:, :d] + dy * dx * patches[..., 1:, 1:])
        if out is not None:
            out.copy_(res)
            return out
        return res
    if mode != 'bilinear':
        return patchify_forward(net, coords, radius)[0]
    _lib.require_cuda(net, coords)
    B, C, H, W = net.shape
    M = coords.shape[1]
    d = 2 * radius + 1
    coords = coords.to(torch.float32).contiguous()
    if out is None:
        out = torch.empty(B, M, C, d, d, dtype=torch.float32, device=net.device)
    else:
        assert tuple(out.shape) == (B, M, C, d, d)
        _lib.require_cuda(out)
    s = out.stride()
    with torch.cuda.device(net.device):
        v = _lib.fmap_view(net)
        _lib.check(_lib.lib().rvo_patchify_bilinear(ctypes.byref(v), _lib.ptr(coords), M, radius,
                                                    _lib.ptr(out), _lib.dtype_code(out), s[0], s[1],
                                                    s[2], s[3], s[4], _lib.stream_ptr()),
                   "rvo_patchify_bilinear")
    return out


def _views(fmap1, fmap2):
    if fmap1.dtype != fmap2.dtype:
        raise RuntimeError("altcorr.corr: fmap1 and fmap2 must have the same dtype")
    return _lib.fmap_view(fmap1), _lib.fmap_view(fmap2)


def corr(fmap1, fmap2, coords, ii, jj, radius=1, dropout=1):
    """ramp.altcorr.corr (correlation.py:71; cuda_corr.forward correlation_kernel.cu:193-233).

    fmap1 [B,Np,C,P,P], fmap2 [B,Nf,C,H2,W2], coords [B,E,2,P,P] f32, ii/jj [E] int64 ->
    [B,E,2R+1,2R+1,P,P] (x-offset dim first) in fmap1's dtype.  fp32 accumulation.  `dropout` only
    affects the reference's backward pass (correlation.py:20-25) and is ignored here.
    """
    if _needs_grad(fmap1, fmap2):
        return _CorrLayer.apply(fmap1, fmap2, coords, ii, jj, radius, dropout)
    _lib.require_cuda(fmap1, fmap2, coords, ii, jj)
    B, E = coords.shape[0], coords.shape[1]
    P = fmap1.shape[-1]
    d = 2 * radius + 1
    coords = coords.to(torch.float32).contiguous()
    ii = ii.to(torch.int64).contiguous()
    jj = jj.to(torch.int64).contiguous()
    out = torch.empty(B, E, d, d, P, P, dtype=fmap1.dtype, device=fmap1.device)
    L = _lib.lib()
    with torch.cuda.device(fmap1.device):
        for b in range(B):
            v1, v2 = _views(fmap1[b], fmap2[b])
            _lib.check(L.rvo_corr_forward(ctypes.byref(v1), ctypes.byref(v2), _lib.ptr(coords[b]),
                                          _lib.ptr(ii), _lib.ptr(jj), E, radius, _lib.ptr(out[b]),
                                          _lib.stream_ptr()), "rvo_corr_forward")
    return out


def corr_pyramid(gmap, pyramid, coords, kk, jj, pmod=0, fmod=0, radius=3, scales=None, out=None):
    """Ramp_vo.corr (ramp/Ramp_vo.py:175-182) in one launch: every pyramid level, the bilinear blend,
    the (x,y) permute and the level stack fused.

    gmap [1,Np,C,P,P]; pyramid: list of [1,Nf,C,H_l,W_l]; coords [1,E,2,P,P]; patch index
    kk % pmod, frame index jj % fmod (0 = no modulo).  Returns [1,E,(2R+1)^2*P*P*len(pyramid)].
    Channels-last fp16 maps (stride(C) == 1) take the tensor-core path.
    """
    _lib.require_cuda(gmap, coords, kk, jj, *pyramid)
    nl = len(pyramid)
    if scales is None:
        scales = [1.0, 0.25, 0.0625][:nl]
    E = coords.shape[1]
    P = gmap.shape[-1]
    d = 2 * radius + 1
    coords = coords.to(torch.float32).contiguous()
    kk = kk.to(torch.int64).contiguous()
    jj = jj.to(torch.int64).contiguous()
    row = d * d * P * P * nl
    if out is None:
        out = torch.empty(1, E, row, dtype=gmap.dtype, device=gmap.device)
    ld = out.stride(-2) if out.numel() else row      # padded rows allowed (pad columns untouched)
    assert out.shape[-1] >= row and out.stride(-1) == 1
    v1 = _lib.fmap_view(gmap[0])
    arr = (_lib.FMap * nl)(*[_lib.fmap_view(p[0]) for p in pyramid])
    for p in pyramid:
        if p.dtype != gmap.dtype:
            raise RuntimeError("altcorr.corr_pyramid: dtype mismatch between gmap and pyramid")
    sc = (ctypes.c_float * nl)(*scales)
    with torch.cuda.device(gmap.device):
        _lib.check(_lib.lib().rvo_corr_pyramid(ctypes.byref(v1), arr, sc, nl, _lib.ptr(coords),
                                               _lib.ptr(kk), _lib.ptr(jj), pmod, fmod, E, radius,
                                               _lib.ptr(out), ld, _lib.stream_ptr()),
                   "rvo_corr_pyramid")
    return out


TILE_GROUP = 56      # halves per (level, pixel) group of the tile layout: 7 rows of 8 (7 used)


def tile_layout_index(nlevels=2, P=3, radius=3):
    """index map from the tile layout of corr_tiles to the reference layout of Ramp_vo.corr:
    ref_index[t] for every used tile-layout column t, and the list of those columns."""
    d = 2 * radius + 1
    cols, ref = [], []
    for lvl in range(nlevels):
        for pix in range(P * P):
            for a in range(d):
                for b in range(d):
                    cols.append((lvl * P * P + pix) * TILE_GROUP + a * 8 + b)
                    ref.append(((b * d + a) * P * P + pix) * nlevels + lvl)
    return torch.tensor(cols), torch.tensor(ref)


def corr_tiles(gmap, pyramid, coords, kk, jj, pmod=0, fmod=0, scales=None, out=None):
    """Ramp_vo.corr on the tcgen05 tensor cores (rvo_corr_tiles): same values as corr_pyramid in the
    TILE layout [1, E, >= 450*levels] (see tile_layout_index); gmap / pyramid must be channels-last
    fp16 with 128 channels, P = 3, radius 3."""
    _lib.require_cuda(gmap, coords, kk, jj, *pyramid)
    nl = len(pyramid)
    if scales is None:
        scales = [1.0, 0.25][:nl]
    E = coords.shape[1]
    coords = coords.to(torch.float32).contiguous()
    kk = kk.to(torch.int64).contiguous()
    jj = jj.to(torch.int64).contiguous()
    row = 9 * nl * TILE_GROUP
    if out is None:
        out = torch.zeros(1, E, (row + 7) // 8 * 8, dtype=torch.float16, device=gmap.device)
    ld = out.stride(-2) if out.numel() else row
    v1 = _lib.fmap_view(gmap[0])
    arr = (_lib.FMap * nl)(*[_lib.fmap_view(p[0]) for p in pyramid])
    sc = (ctypes.c_float * nl)(*scales)
    L = _lib.lib()
    nb = L.rvo_corr_tiles_ws_bytes(arr, nl, E)
    if nb < 0:
        raise RuntimeError("altcorr.corr_tiles: %s" % L.rvo_last_error().decode())
    ws = _lib.Workspace.get(gmap.device, nb, "corr_tiles")
    with torch.cuda.device(gmap.device):
        _lib.check(L.rvo_corr_tiles(ctypes.byref(v1), arr, sc, nl, _lib.ptr(coords), _lib.ptr(kk),
                                    _lib.ptr(jj), pmod, fmod, E, _lib.ptr(out), ld, _lib.ptr(ws),
                                    ws.numel(), _lib.stream_ptr()), "rvo_corr_tiles")
    return out
