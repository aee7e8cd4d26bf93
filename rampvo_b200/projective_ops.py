"""Drop-in for ramp.projective_ops (ramp/projective_ops.py:5-118): the ~15 small torch ops + 3
lietorch kernels of one `transform` call are a single fused kernel here (rvo_transform*)."""
import torch

from . import _lib
from .lietorch import SE3

MIN_DEPTH = 0.2


def extract_intrinsics(intrinsics):
    return intrinsics[..., None, None, :].unbind(dim=-1)


def coords_grid(ht, wd, **kwargs):
    y, x = torch.meshgrid(torch.arange(ht).to(**kwargs).float(),
                          torch.arange(wd).to(**kwargs).float(), indexing="ij")
    return torch.stack([x, y], dim=-1)


def _prep(poses, patches, intrinsics, *idx):
    pd = getattr(poses, "data", poses)
    _lib.require_cuda(pd, patches, intrinsics, *idx)
    if pd.shape[0] != 1 or patches.shape[0] != 1:
        raise RuntimeError("projective_ops: batch size must be 1 (ramp/utils.py:238)")
    P = patches.shape[-1]
    pv = pd.to(torch.float32).contiguous().view(-1, 7)
    qv = patches.to(torch.float32).contiguous().view(-1, 3, P, P)
    kv = intrinsics.to(torch.float32).contiguous().view(-1, 4)
    idx = [t.to(torch.int64).contiguous() for t in idx]
    return pv, qv, kv, P, idx


def transform(poses, patches, intrinsics, ii, jj, kk, depth=False, valid=False, jacobian=False,
              tonly=False):
    """pops.transform (projective_ops.py:50-101): reproject patch kk from frame ii into frame jj.

    Returns x1 [1,E,P,P,2] (or [...,3] with depth=True); with valid=True also (Z > 0.2) [1,E,P,P];
    with jacobian=True (x1, valid [1,E], (Ji [1,E,2,6], Jj [1,E,2,6], Jz [1,E,2,1]))."""
    pd = getattr(poses, "data", poses)
    if torch.is_grad_enabled() and (pd.requires_grad or patches.requires_grad):
        return _transform_autograd(poses, patches, intrinsics, ii, jj, kk, depth, valid, jacobian, tonly)
    pv, qv, kv, P, (ii, jj, kk) = _prep(poses, patches, intrinsics, ii, jj, kk)
    E = ii.numel()
    dev = pv.device
    L = _lib.lib()
    x1 = torch.empty(1, E, P, P, 2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        if jacobian:
            v = torch.empty(1, E, dtype=torch.float32, device=dev)
            Ji = torch.empty(1, E, 2, 6, dtype=torch.float32, device=dev)
            Jj = torch.empty(1, E, 2, 6, dtype=torch.float32, device=dev)
            Jz = torch.empty(1, E, 2, 1, dtype=torch.float32, device=dev)
            _lib.check(L.rvo_transform_jac(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(kv), _lib.ptr(ii),
                                           _lib.ptr(jj), _lib.ptr(kk), E, P, _lib.ptr(x1),
                                           _lib.ptr(v), _lib.ptr(Ji), _lib.ptr(Jj), _lib.ptr(Jz),
                                           _lib.stream_ptr()), "rvo_transform_jac")
            return x1, v, (Ji, Jj, Jz)
        d = torch.empty(1, E, P, P, dtype=torch.float32, device=dev) if depth else None
        v = torch.empty(1, E, P, P, dtype=torch.float32, device=dev) if valid else None
        _lib.check(L.rvo_transform(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(kv), _lib.ptr(ii),
                                   _lib.ptr(jj), _lib.ptr(kk), E, P,
                                   _lib.RVO_TF_TONLY if tonly else 0, _lib.ptr(x1), None,
                                   _lib.ptr(d), _lib.ptr(v), _lib.stream_ptr()), "rvo_transform")
    if depth:
        x1 = torch.cat([x1, d[..., None]], dim=-1)
    if valid:
        return x1, v
    return x1


def _transform_autograd(poses, patches, intrinsics, ii, jj, kk, depth, valid, jacobian, tonly):
    """The same map as the fused kernel written with tensor ops, used ONLY while gradients are being recorded
    (training unroll, ramp/net.py:341-371): autograd differentiates it w.r.t. poses and patches."""
    poses = poses if isinstance(poses, SE3) else SE3(poses)
    X0 = iproj(patches[:, kk], intrinsics[:, ii])
    Gij = poses[:, jj] * poses[:, ii].inv()
    if tonly:
        ident = torch.zeros_like(Gij.data[..., 3:])
        ident[..., 3] = 1.0
        Gij = SE3(torch.cat([Gij.data[..., :3], ident], dim=-1))
    X1 = Gij[:, :, None, None] * X0
    x1 = proj(X1, intrinsics[:, jj], depth)
    if jacobian:
        p = X1.shape[2]
        X, Y, Z, H = X1[..., p // 2, p // 2, :].unbind(dim=-1)
        o = torch.zeros_like(H)
        fx, fy, cx, cy = intrinsics[:, jj].unbind(dim=-1)
        d = torch.where(Z.abs() > 0.2, 1.0 / torch.where(Z.abs() > 0.2, Z, torch.ones_like(Z)), o)
        Ja = torch.stack([H, o, o, o, Z, -Y,
                          o, H, o, -Z, o, X,
                          o, o, H, Y, -X, o,
                          o, o, o, o, o, o], dim=-1).view(1, len(ii), 4, 6)
        Jp = torch.stack([fx * d, o, -fx * X * d * d, o,
                          o, fy * d, -fy * Y * d * d, o], dim=-1).view(1, len(ii), 2, 4)
        Jj = torch.matmul(Jp, Ja)
        Ji = -Gij[:, :, None].adjT(Jj)
        Jz = torch.matmul(Jp, Gij.matrix()[..., :, 3:])
        return x1, (Z > 0.2).float(), (Ji, Jj, Jz)
    if valid:
        return x1, (X1[..., 2] > 0.2).float()
    return x1


def reproject_cf(poses, patches, intrinsics, ii, jj, kk, out=None):
    """Ramp_vo.reproject (ramp/Ramp_vo.py:184-192): transform + permute(0,1,4,2,3) fused; returns
    coords [1,E,2,P,P]."""
    pv, qv, kv, P, (ii, jj, kk) = _prep(poses, patches, intrinsics, ii, jj, kk)
    E = ii.numel()
    if out is None:
        out = torch.empty(1, E, 2, P, P, dtype=torch.float32, device=pv.device)
    with torch.cuda.device(pv.device):
        _lib.check(_lib.lib().rvo_transform(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(kv), _lib.ptr(ii),
                                            _lib.ptr(jj), _lib.ptr(kk), E, P, 0, None,
                                            _lib.ptr(out), None, None, _lib.stream_ptr()),
                   "rvo_transform")
    return out


def iproj(patches, intrinsics):
    """projective_ops.py:16-26 (tensor plumbing; not on the fused path)."""
    x, y, d = patches.unbind(dim=2)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    i = torch.ones_like(d)
    return torch.stack([(x - cx) / fx, (y - cy) / fy, i, d], dim=-1)


def proj(X, intrinsics, depth=False):
    """projective_ops.py:29-47."""
    X, Y, Z, W = X.unbind(dim=-1)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    d = 1.0 / Z.clamp(min=0.1)
    x = fx * (d * X) + cx
    y = fy * (d * Y) + cy
    if depth:
        return torch.stack([x, y, d], dim=-1)
    return torch.stack([x, y], dim=-1)


def point_cloud(poses, patches, intrinsics, ix):
    """pops.point_cloud (projective_ops.py:103-105): [1,m,P,P,4] homogeneous points."""
    return SE3(getattr(poses, "data", poses))[:, ix, None, None].inv() * iproj(patches, intrinsics[:, ix])


def point_cloud_centers(poses, patches, intrinsics, ix):
    """The part of point_cloud Ramp_vo.update keeps (ramp/Ramp_vo.py:308-310): centre pixel,
    X/W normalised -> [m,3], one fused kernel."""
    pv, qv, kv, P, (ix,) = _prep(poses, patches, intrinsics, ix)
    m = ix.numel()
    out = torch.empty(m, 3, dtype=torch.float32, device=pv.device)
    with torch.cuda.device(pv.device):
        _lib.check(_lib.lib().rvo_point_cloud(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(kv), _lib.ptr(ix),
                                              m, P, _lib.ptr(out), _lib.stream_ptr()),
                   "rvo_point_cloud")
    return out


def flow_mag(poses, patches, intrinsics, ii, jj, kk, beta=0.3):
    """pops.flow_mag (projective_ops.py:108-118): [1,E,P,P], the three transforms fused."""
    pv, qv, kv, P, (ii, jj, kk) = _prep(poses, patches, intrinsics, ii, jj, kk)
    E = ii.numel()
    out = torch.empty(1, E, P, P, dtype=torch.float32, device=pv.device)
    with torch.cuda.device(pv.device):
        _lib.check(_lib.lib().rvo_flow_mag(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(kv), _lib.ptr(ii),
                                           _lib.ptr(jj), _lib.ptr(kk), E, P, float(beta),
                                           _lib.ptr(out), _lib.stream_ptr()), "rvo_flow_mag")
    return out
