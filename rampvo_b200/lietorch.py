"""SE3 type layer: the subset of ramp.lietorch the hot path touches (ramp/lietorch/groups.py:51-322).

The reference's `lietorch_backends` extension cannot be built without Eigen and
the north star keeps "lietorch SE(3) types" as host plumbing, so this module re-provides the SE3
forward operators {exp, log, inv, mul, act, adjT, matrix, retr} on top of plain tensor arithmetic,
following ramp/lietorch/include/so3.h:55-60,115-215 and se3.h:36-142.  It only ever handles a few
poses at a time (motion model, keyframe bookkeeping); the per-edge geometry of the hot loop runs in
the fused kernels of rampvo_b200.projective_ops / fastba, never through this class.

Data layout [tx,ty,tz,qx,qy,qz,qw] (groups.py:273); quaternions are re-normalised on load like the
reference constructors do (so3.h:31-37).
"""
import math

import numpy as np
import torch

EPS = 1e-6  # ramp/lietorch/include/common.h:7


def _cross(a, b):
    return torch.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                        a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                        a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], dim=-1)


def _normalize_q(q):
    return q / q.norm(dim=-1, keepdim=True)


def _rot(q, p):
    """so3.h:55-60"""
    qv, qw = q[..., :3], q[..., 3:4]
    uv = _cross(qv, p)
    uv = uv + uv
    return p + qw * uv + _cross(qv, uv)


def _qmul(a, b):
    """Hamilton product, xyzw."""
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw,
                        aw * bw - ax * bx - ay * by - az * bz], dim=-1)


def _hat(phi):
    o = torch.zeros_like(phi[..., 0])
    x, y, z = phi.unbind(-1)
    return torch.stack([o, -z, y, z, o, -x, -y, x, o], dim=-1).view(phi.shape[:-1] + (3, 3))


def _so3_exp(phi):
    """so3.h:153-170"""
    theta2 = (phi * phi).sum(-1, keepdim=True)
    small = theta2 < EPS * EPS
    # the square root is taken of a value bounded away from zero so that autograd never sees d sqrt(0) = inf
    # (the training unroll differentiates through exp / log of identity-like poses); values are unchanged
    theta = torch.where(small, torch.ones_like(theta2), theta2).sqrt()
    theta4 = theta2 * theta2
    safe = theta
    imag = torch.where(small, 0.5 - (1.0 / 48.0) * theta2 + (1.0 / 3840.0) * theta4,
                       torch.sin(0.5 * safe) / safe)
    real = torch.where(small, 1.0 - (1.0 / 8.0) * theta2 + (1.0 / 384.0) * theta4,
                       torch.cos(0.5 * safe))
    return _normalize_q(torch.cat([imag * phi, real], dim=-1))


def _so3_log(q):
    """so3.h:115-151"""
    qv, w = q[..., :3], q[..., 3:4]
    sq = (qv * qv).sum(-1, keepdim=True)
    small = sq < EPS * EPS
    n = torch.where(small, torch.ones_like(sq), sq).sqrt()
    safe_n = n
    wz = w.abs() < EPS
    safe_w = torch.where(wz, torch.ones_like(w), w)
    f_small = 2.0 / safe_w - (2.0 / 3.0) * sq / (safe_w * safe_w * safe_w)
    f_wz = torch.where(w > 0, math.pi / safe_n, -math.pi / safe_n)
    f_gen = 2.0 * torch.atan(safe_n / safe_w) / safe_n
    f = torch.where(small, f_small, torch.where(wz, f_wz, f_gen))
    return f * qv


def _left_jacobian(phi):
    """so3.h:172-191"""
    I = torch.eye(3, dtype=phi.dtype, device=phi.device).expand(phi.shape[:-1] + (3, 3))
    Phi = _hat(phi)
    Phi2 = Phi @ Phi
    theta2 = (phi * phi).sum(-1, keepdim=True)
    small = theta2 < EPS * EPS
    st2 = torch.where(small, torch.ones_like(theta2), theta2)
    st = st2.sqrt()
    c1 = torch.where(small, 0.5 - (1.0 / 24.0) * theta2, (1.0 - torch.cos(st)) / st2)
    c2 = torch.where(small, 1.0 / 6.0 - (1.0 / 120.0) * theta2, (st - torch.sin(st)) / (st2 * st))
    return I + c1[..., None] * Phi + c2[..., None] * Phi2


def _left_jacobian_inverse(phi):
    """so3.h:193-208"""
    I = torch.eye(3, dtype=phi.dtype, device=phi.device).expand(phi.shape[:-1] + (3, 3))
    Phi = _hat(phi)
    Phi2 = Phi @ Phi
    theta2 = (phi * phi).sum(-1, keepdim=True)
    small = theta2 < EPS * EPS
    st = torch.where(small, torch.ones_like(theta2), theta2).sqrt()
    half = 0.5 * st
    c2 = torch.where(small, torch.full_like(theta2, 1.0 / 12.0),
                     (1.0 - st * torch.cos(half) / (2.0 * torch.sin(half))) / (st * st))
    return I - 0.5 * Phi + c2[..., None] * Phi2


class SE3:
    group_name = 'SE3'
    group_id = 3
    manifold_dim = 6
    embedded_dim = 7
    id_elem = torch.as_tensor([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0])

    def __init__(self, data):
        if isinstance(data, SE3):
            data = data.data
        self.data = data

    def __repr__(self):
        return "{}: size={}, device={}, dtype={}".format(self.group_name, self.shape, self.device,
                                                         self.dtype)

    @property
    def shape(self):
        return self.data.shape[:-1]

    @property
    def device(self):
        return self.data.device

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def tangent_shape(self):
        return self.data.shape[:-1] + (self.manifold_dim,)

    # -- constructors (groups.py:80-128)
    @classmethod
    def Identity(cls, *batch_shape, **kwargs):
        if isinstance(batch_shape[0], (tuple, list, torch.Size)):
            batch_shape = tuple(batch_shape[0])
        numel = int(np.prod(batch_shape))
        data = cls.id_elem.reshape(1, -1)
        if 'device' in kwargs:
            data = data.to(kwargs['device'])
        if 'dtype' in kwargs:
            data = data.type(kwargs['dtype'])
        data = data.repeat(numel, 1)
        return cls(data).view(tuple(batch_shape))

    @classmethod
    def IdentityLike(cls, G):
        return cls.Identity(G.shape, device=G.data.device, dtype=G.data.dtype)

    @classmethod
    def Random(cls, *batch_shape, sigma=1.0, **kwargs):
        if isinstance(batch_shape[0], (tuple, list)):
            batch_shape = tuple(batch_shape[0])
        xi = torch.randn(tuple(batch_shape) + (cls.manifold_dim,), **kwargs)
        return cls.exp(sigma * xi)

    # -- split helpers
    def _tq(self):
        return self.data[..., :3], _normalize_q(self.data[..., 3:7])

    # -- group operators
    @classmethod
    def exp(cls, x):
        """se3.h:134-142"""
        tau, phi = x[..., :3], x[..., 3:]
        q = _so3_exp(phi)
        t = (_left_jacobian(phi) @ tau[..., None])[..., 0]
        return cls(torch.cat([t, q], dim=-1))

    def log(self):
        """se3.h:124-132"""
        t, q = self._tq()
        phi = _so3_log(q)
        tau = (_left_jacobian_inverse(phi) @ t[..., None])[..., 0]
        return torch.cat([tau, phi], dim=-1)

    def inv(self):
        """se3.h:36-38"""
        t, q = self._tq()
        qi = _normalize_q(q * torch.as_tensor([-1.0, -1.0, -1.0, 1.0], dtype=q.dtype, device=q.device))
        return SE3(torch.cat([-_rot(qi, t), qi], dim=-1))

    def mul(self, other):
        """se3.h:45-47"""
        t1, q1 = self._tq()
        t2, q2 = other._tq()
        q = _normalize_q(_qmul(q1, q2))
        t = t1 + _rot(q1, t2)
        t, q = torch.broadcast_tensors(t, q[..., :3])[0], q
        return SE3(torch.cat([t, q.expand(t.shape[:-1] + (4,))], dim=-1))

    def retr(self, a):
        """groups.py:153-156: Exp(a) * X"""
        return SE3.exp(a).mul(self)

    def act(self, p):
        """se3.h:49-56"""
        t, q = self._tq()
        if p.shape[-1] == 3:
            return _rot(q, p) + t
        return torch.cat([_rot(q, p[..., :3]) + t * p[..., 3:4], p[..., 3:4]], dim=-1)

    def adjT(self, a):
        """se3.h:84-86 (Adj^T a): [R^T a_t, R^T a_w + R^T (a_t x t)]"""
        t, q = self._tq()
        qi = q * torch.as_tensor([-1.0, -1.0, -1.0, 1.0], dtype=q.dtype, device=q.device)
        at, aw = a[..., :3], a[..., 3:]
        return torch.cat([_rot(qi, at), _rot(qi, aw) + _rot(qi, _cross(at, t))], dim=-1)

    def matrix(self):
        """groups.py:183-187"""
        t, q = self._tq()
        I = torch.eye(3, dtype=self.dtype, device=self.device)
        I = I.view([1] * (self.data.dim() - 1) + [3, 3])
        R = _rot(q[..., None, :], I).transpose(-1, -2)  # columns = R e_k
        top = torch.cat([R, t[..., None]], dim=-1)
        bot = torch.zeros_like(top[..., :1, :])
        bot[..., 0, 3] = 1.0
        return torch.cat([top, bot], dim=-2)

    def translation(self):
        t = self.data[..., :3]
        return torch.cat([t, torch.ones_like(t[..., :1])], dim=-1)

    def scale(self, s):
        t, q = self.data.split([3, 4], -1)
        return SE3(torch.cat([t * s.unsqueeze(-1), q], dim=-1))

    # -- tensor plumbing (groups.py:189-232)
    def detach(self):
        return SE3(self.data.detach())

    def view(self, dims):
        return SE3(self.data.view(tuple(dims) + (self.embedded_dim,)))

    def __mul__(self, other):
        if isinstance(other, SE3):
            return self.mul(other)
        if isinstance(other, torch.Tensor):
            return self.act(other)
        return NotImplemented

    def __getitem__(self, index):
        return SE3(self.data[index])

    def __setitem__(self, index, item):
        self.data[index] = item.data

    def to(self, *args, **kwargs):
        return SE3(self.data.to(*args, **kwargs))

    def cpu(self):
        return SE3(self.data.cpu())

    def cuda(self):
        return SE3(self.data.cuda())

    def float(self, device=None):
        return SE3(self.data.float())

    def double(self, device=None):
        return SE3(self.data.double())

    def unbind(self, dim=0):
        return [SE3(x) for x in self.data.unbind(dim=dim)]


def cat(group_list, dim):
    return SE3(torch.cat([X.data for X in group_list], dim=dim))


def stack(group_list, dim):
    return SE3(torch.stack([X.data for X in group_list], dim=dim))
