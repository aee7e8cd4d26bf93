"""`lietorch_backends` stand-in (TEST / BASELINE INFRASTRUCTURE — never imported by rampvo_b200/).

The reference's Lie-group extension (ramp/lietorch/src/lietorch.cpp:286-314) cannot be built here: every
TU includes Eigen 3.4.0, which is neither vendored nor in the image.  This module
re-provides the forward entry points the reference's own Python wrappers (ramp/lietorch/group_ops.py:28-66)
bind, for SO3 (group id 1) and SE3 (group id 3), as batched tensor arithmetic on [n, dim] inputs, so that
ramp/lietorch/groups.py, ramp/projective_ops.py, ramp/ba.py and ramp/Ramp_vo.py run UNMODIFIED on top of
it, on CPU or GPU.  Formulas: ramp/lietorch/include/so3.h:55-60 (rotate), :115-151 (Log), :153-170 (Exp),
:172-208 (left Jacobian and inverse); se3.h:36-38 (inv), :45-47 (mul), :53-56 (act4), :58-67,84-86
(Adj / AdjT), :124-142 (Log / Exp).  Backward entry points raise: only inference is driven through it.
"""
import math

import torch

EPS = 1e-6          # ramp/lietorch/include/common.h:7
SO3, SE3 = 1, 3


def _cross(a, b):
    return torch.cross(a, b, dim=-1)


def _hat(v):
    o = torch.zeros_like(v[:, 0])
    x, y, z = v.unbind(-1)
    return torch.stack([o, -z, y, z, o, -x, -y, x, o], -1).view(-1, 3, 3)


def _rot(q, p):
    qv, qw = q[:, :3], q[:, 3:4]
    uv = 2 * _cross(qv, p)
    return p + qw * uv + _cross(qv, uv)


def _qmul(a, b):
    av, aw, bv, bw = a[:, :3], a[:, 3:], b[:, :3], b[:, 3:]
    return torch.cat([aw * bv + bw * av + _cross(av, bv), aw * bw - (av * bv).sum(-1, keepdim=True)], -1)


def _qinv(q):
    return torch.cat([-q[:, :3], q[:, 3:]], -1)


def _qnorm(q):
    return q / q.norm(dim=-1, keepdim=True)


def _rotmat(q):
    I = torch.eye(3, dtype=q.dtype, device=q.device)
    return torch.stack([_rot(q, I[k].expand(q.shape[0], 3)) for k in range(3)], -1)


def _so3_exp(phi):
    t2 = (phi * phi).sum(-1, keepdim=True)
    t = t2.sqrt()
    small = t < EPS
    ts = torch.where(small, torch.ones_like(t), t)
    imag = torch.where(small, 0.5 - t2 / 48 + t2 * t2 / 3840, torch.sin(0.5 * ts) / ts)
    real = torch.where(small, 1 - t2 / 8 + t2 * t2 / 384, torch.cos(0.5 * ts))
    return _qnorm(torch.cat([imag * phi, real], -1))


def _so3_log(q):
    qv, w = q[:, :3], q[:, 3:]
    sq = (qv * qv).sum(-1, keepdim=True)
    n = sq.sqrt()
    small = sq < EPS * EPS
    ns = torch.where(small, torch.ones_like(n), n)
    wz = w.abs() < EPS
    ws = torch.where(wz, torch.ones_like(w), w)
    f = torch.where(small, 2 / ws - (2.0 / 3) * sq / (ws * ws * ws),
                    torch.where(wz, torch.where(w > 0, math.pi / ns, -math.pi / ns),
                                2 * torch.atan(ns / ws) / ns))
    return f * qv


def _left_jacobian(phi):
    I = torch.eye(3, dtype=phi.dtype, device=phi.device)
    Phi = _hat(phi)
    t2 = (phi * phi).sum(-1, keepdim=True)
    t = t2.sqrt()
    small = t < EPS
    t2s, ts = torch.where(small, torch.ones_like(t2), t2), torch.where(small, torch.ones_like(t), t)
    c1 = torch.where(small, 0.5 - t2 / 24, (1 - torch.cos(ts)) / t2s)
    c2 = torch.where(small, 1.0 / 6 - t2 / 120, (ts - torch.sin(ts)) / (t2s * ts))
    return I + c1[..., None] * Phi + c2[..., None] * (Phi @ Phi)


def _left_jacobian_inverse(phi):
    I = torch.eye(3, dtype=phi.dtype, device=phi.device)
    Phi = _hat(phi)
    t = (phi * phi).sum(-1, keepdim=True).sqrt()
    small = t < EPS
    ts = torch.where(small, torch.ones_like(t), t)
    c2 = torch.where(small, torch.full_like(t, 1.0 / 12),
                     (1 - ts * torch.cos(0.5 * ts) / (2 * torch.sin(0.5 * ts))) / (ts * ts))
    return I - 0.5 * Phi + c2[..., None] * (Phi @ Phi)


def _split(gid, X):
    """-> (t or None, normalised q); the reference constructors renormalise on load (so3.h:31-37)"""
    if gid == SO3:
        return None, _qnorm(X[:, :4])
    if gid == SE3:
        return X[:, :3], _qnorm(X[:, 3:7])
    raise NotImplementedError("lietorch_backends stand-in: group id %d (SO3 = 1 and SE3 = 3 only)" % gid)


def _join(t, q):
    return q if t is None else torch.cat([t, q], -1)


def expm(gid, a):
    if gid == SO3:
        return _so3_exp(a)
    _split(gid, torch.zeros(1, 7))
    tau, phi = a[:, :3], a[:, 3:6]
    return torch.cat([(_left_jacobian(phi) @ tau[..., None])[..., 0], _so3_exp(phi)], -1)


def logm(gid, X):
    t, q = _split(gid, X)
    phi = _so3_log(q)
    if t is None:
        return phi
    return torch.cat([(_left_jacobian_inverse(phi) @ t[..., None])[..., 0], phi], -1)


def inv(gid, X):
    t, q = _split(gid, X)
    qi = _qinv(q)
    return _join(None if t is None else -_rot(qi, t), qi)


def mul(gid, X, Y):
    tx, qx = _split(gid, X)
    ty, qy = _split(gid, Y)
    return _join(None if tx is None else tx + _rot(qx, ty), _qmul(qx, qy))


def _adj_matrix(gid, X):
    t, q = _split(gid, X)
    R = _rotmat(q)
    if t is None:
        return R
    Z = torch.zeros_like(R)
    return torch.cat([torch.cat([R, _hat(t) @ R], -1), torch.cat([Z, R], -1)], -2)


def adj(gid, X, a):
    return (_adj_matrix(gid, X) @ a[..., None])[..., 0]


def adjT(gid, X, a):
    return (_adj_matrix(gid, X).transpose(-1, -2) @ a[..., None])[..., 0]


def act(gid, X, p):
    t, q = _split(gid, X)
    y = _rot(q, p)
    return y if t is None else y + t


def act4(gid, X, p):
    t, q = _split(gid, X)
    y = _rot(q, p[:, :3])
    if t is not None:
        y = y + t * p[:, 3:4]
    return torch.cat([y, p[:, 3:4]], -1)


def as_matrix(gid, X):
    t, q = _split(gid, X)
    n = X.shape[0]
    T = torch.eye(4, dtype=X.dtype, device=X.device).repeat(n, 1, 1)
    T[:, :3, :3] = _rotmat(q)
    if t is not None:
        T[:, :3, 3] = t
    return T


def _no_backward(name):
    def f(*a, **k):
        raise NotImplementedError("lietorch_backends stand-in: %s (inference only)" % name)
    return f


for _n in ("expm", "logm", "inv", "mul", "adj", "adjT", "act", "act4"):
    globals()[_n + "_backward"] = _no_backward(_n + "_backward")
Jinv = _no_backward("Jinv")
projector = _no_backward("projector")
