"""Drop-in for ramp.ba (ramp/ba.py:86-182): the differentiable Gauss-Newton bundle-adjustment step of the TRAINING
unroll (ramp/net.py:352-367).  Inference uses the fused CUDA solver (rampvo_b200.fastba); this module is plain
tensor algebra so that autograd carries the loss back into the update operator's targets and weights.

Same semantics as the reference: residual gate 250 px, caller-given bounds, damping `ep` + relative `lm`, depth
clamp [1e-3, 10], first `fixedp` poses fixed, a failed Cholesky yields a zero step (ba.py:16-20).  Organised
differently: every edge contributes one 13-column row block [Ji | Jj | Jz]; the normal equations are accumulated
with three index_add_ calls on flat (pose, pose) / (pose, patch) / patch keys instead of block 5-D tensors and
torch_scatter.
"""
import torch

from . import projective_ops as pops
from .lietorch import SE3


class CholeskySolver(torch.autograd.Function):
    """ba.py:12-37: solve H x = b by Cholesky; zero step (and no gradient) when H is not positive definite"""

    @staticmethod
    def forward(ctx, H, b):
        U, info = torch.linalg.cholesky_ex(H)
        if torch.any(info):
            ctx.failed = True
            return torch.zeros_like(b)
        xs = torch.cholesky_solve(b, U)
        ctx.save_for_backward(U, xs)
        ctx.failed = False
        return xs

    @staticmethod
    def backward(ctx, grad_x):
        if ctx.failed:
            return None, None
        U, xs = ctx.saved_tensors
        dz = torch.cholesky_solve(grad_x, U)
        return -torch.matmul(xs, dz.transpose(-1, -2)), dz


def _accumulate(blocks, key, size):
    """sum `blocks` [E, ...] into `size` slots by `key` [E]; negative keys (fixed poses) are dropped"""
    keep = (key >= 0) & (key < size)
    out = torch.zeros((size,) + tuple(blocks.shape[1:]), dtype=blocks.dtype, device=blocks.device)
    return out.index_add_(0, key[keep], blocks[keep])


def BA(poses, patches, intrinsics, targets, weights, lmbda, ii, jj, kk, bounds, ep=100.0, PRINT=False, fixedp=1,
       structure_only=False):
    """one damped Gauss-Newton step on poses [1,N,7] (SE3) and the inverse depths of patches [1,K,3,P,P];
    returns (poses, patches) like the reference"""
    poses = poses if isinstance(poses, SE3) else SE3(poses)
    n_all = int(max(ii.max().item(), jj.max().item())) + 1
    coords, valid, (Ji, Jj, Jz) = pops.transform(poses, patches, intrinsics, ii, jj, kk, jacobian=True)
    p = coords.shape[3]
    centre = coords[..., p // 2, p // 2, :]
    r = targets - centre
    valid = valid * (r.norm(dim=-1) < 250).float()
    inside = ((centre[..., 0] > bounds[0]) & (centre[..., 1] > bounds[1]) &
              (centre[..., 0] < bounds[2]) & (centre[..., 1] < bounds[3]))
    valid = valid * inside.float()
    if PRINT:
        print((r * valid[..., None]).norm(dim=-1).mean().item())
    r = (valid[..., None] * r)[0]                              # [E,2]
    w = (valid[..., None] * weights)[0]                        # [E,2]
    Ji, Jj, Jz = Ji[0], Jj[0], Jz[0]                           # [E,2,6] [E,2,6] [E,2,1]

    n = n_all - fixedp                                         # free poses
    pi, pj = ii - fixedp, jj - fixedp
    kx, kc = torch.unique(kk, return_inverse=True, sorted=True)
    m = len(kx)
    wJi, wJj, wJz = w[..., None] * Ji, w[..., None] * Jj, w[..., None] * Jz
    C = _accumulate((wJz * Jz).sum(dim=(1, 2)), kc, m)         # [m]
    u = _accumulate((wJz[..., 0] * r).sum(dim=1), kc, m)       # [m]
    lm_t = lmbda.reshape(-1) if isinstance(lmbda, torch.Tensor) else lmbda
    Q = 1.0 / (C + lm_t)

    if structure_only or n <= 0:
        dZ = Q * u
        dX = None
    else:
        tr = lambda a: a.transpose(1, 2)
        pair = lambda a, b: torch.where((a >= 0) & (b >= 0), a * n + b, torch.full_like(a, -1))
        # pose-pose blocks on the flat key i*n + j, pose-patch blocks on i*m + k
        B = (_accumulate(tr(wJi) @ Ji, pair(pi, pi), n * n) + _accumulate(tr(wJi) @ Jj, pair(pi, pj), n * n) +
             _accumulate(tr(wJj) @ Ji, pair(pj, pi), n * n) + _accumulate(tr(wJj) @ Jj, pair(pj, pj), n * n))
        B = B.view(n, n, 6, 6).permute(0, 2, 1, 3).reshape(6 * n, 6 * n)
        pk = lambda a: torch.where(a >= 0, a * m + kc, torch.full_like(a, -1))
        E = (_accumulate((tr(wJi) @ Jz)[..., 0], pk(pi), n * m) + _accumulate((tr(wJj) @ Jz)[..., 0], pk(pj), n * m))
        E = E.view(n, m, 6).permute(0, 2, 1).reshape(6 * n, m)                       # [6n, m]
        v = (_accumulate((tr(wJi) @ r[..., None])[..., 0], pi, n) +
             _accumulate((tr(wJj) @ r[..., None])[..., 0], pj, n)).reshape(6 * n, 1)
        EQ = E * Q[None, :]
        S = B - EQ @ E.t()
        y = v - EQ @ u[:, None]
        S = S + (ep + 1e-4 * S) * torch.eye(6 * n, dtype=S.dtype, device=S.device)   # block_solve damping, ba.py:71
        dX = CholeskySolver.apply(S[None], y[None])[0]                              # [6n,1]
        dZ = Q * (u - (E.t() @ dX)[:, 0])
        dX = dX.view(n, 6)

    x, yv, disps = patches.unbind(dim=2)
    upd = torch.zeros(disps.shape[1], dtype=disps.dtype, device=disps.device).index_add_(0, kx, dZ)
    disps = (disps + upd.view(1, -1, 1, 1)).clamp(min=1e-3, max=10.0)
    patches = torch.stack([x, yv, disps], dim=2)
    if dX is not None:
        step = torch.zeros(poses.data.shape[1], 6, dtype=dX.dtype, device=dX.device)
        step[fixedp:fixedp + n] = dX
        poses = poses.retr(step[None])
    return poses, patches
