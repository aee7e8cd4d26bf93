"""Patch-graph sharding across the GPUs of one box.

The active patch graph is partitioned BY SOURCE FRAME: every edge of a patch, every SoftAgg group
(keyed by kk or by ii*12345+jj) and every neighbour chain then live on one rank, so reproject, corr,
the update operator, the patch blocks C/u/E and the depth updates are rank-local.  The ONE exchange
per Gauss-Newton iteration is the all-reduce (sum) of the reduced camera system [S | y]
(6N x (6N+1) fp32; 14.6 KB at default.yaml, 130 KB at precise.yaml) over NCCL / NVLink; every rank
then runs the identical dense solve and retracts its replica of the poses — no broadcast.
The reference has no multi-GPU support at all.
"""
import torch
import torch.distributed as dist

from . import _lib


def owner_of_frame(frames, world):
    """rank that owns the patches (and all their edges) of source frame i: round-robin i % world"""
    return frames % world


def partition_edges(ii, world, rank):
    """boolean mask of the edges this rank owns"""
    return owner_of_frame(ii, world) == rank


def reduce_system(Sy, group=None):
    """sum [S | y] over ranks, in place (NCCL on CUDA tensors, gloo on CPU tensors)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(Sy, op=dist.ReduceOp.SUM, group=group)
    return Sy


def exchange_depths(patches, frame_lo, frame_hi, M, group=None):
    """make every rank's replica of the inverse depths of frames [frame_lo, frame_hi) current:
    each rank contributes the depths of the frames it owns, the rest as zeros, summed over ranks.
    patches: [n_patches, 3, P, P] fp32 (device or CPU)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return patches
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sl = patches[frame_lo * M:frame_hi * M, 2]
    frames = torch.arange(frame_lo, frame_hi, device=patches.device).repeat_interleave(M)
    mine = (owner_of_frame(frames, world) == rank).view(-1, 1, 1).to(sl.dtype)
    buf = (sl * mine).contiguous()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    sl.copy_(buf)
    return patches


def sharded_BA(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations=2,
               group=None):
    """fastba.BA over a sharded graph.  poses [1,n,7], patches [1,K,3,P,P], intrinsics [1,n,4] are
    replicated; target / weight / ii / jj / kk hold ONLY this rank's edges.  Poses end up identical on
    every rank; depths are updated for the local patches (call exchange_depths to replicate them)."""
    _lib.require_cuda(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk)
    L = _lib.lib()
    P = patches.shape[-1]
    pv, qv, kv = poses.view(-1, 7), patches.view(-1, 3, P, P), intrinsics.view(-1, 4)
    tv = target.to(torch.float32).contiguous().view(-1, 2)
    wv = weight.to(torch.float32).contiguous().view(-1, 2)
    lm = lmbda.to(torch.float32).contiguous().view(-1)
    ii, jj, kk = [t.to(torch.int64).contiguous() for t in (ii, jj, kk)]
    E, N = ii.numel(), t1 - t0
    n6 = 6 * N
    dev = pv.device
    # every rank needs a workspace (and an [S|y]) even when it owns no edge of this window
    Ews = max(E, 1)
    ws = _lib.Workspace.get(dev, L.rvo_ba_ws_bytes(Ews, qv.shape[0], N), "ba_sharded")
    Sy = torch.zeros(max(n6, 1), n6 + 1, dtype=torch.float32, device=dev)
    st = _lib.stream_ptr(dev)
    with torch.cuda.device(dev):
        if E:
            _lib.check(L.rvo_ba_plan(_lib.ptr(kk), _lib.ptr(jj), E, pv.shape[0], qv.shape[0], N,
                                     _lib.ptr(ws), ws.numel(), st), "rvo_ba_plan")
        for _ in range(iterations):
            if E:
                _lib.check(L.rvo_ba_assemble(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(kv), _lib.ptr(tv),
                                             _lib.ptr(wv), _lib.ptr(lm), _lib.ptr(ii), _lib.ptr(jj), E,
                                             qv.shape[0], P, t0, t1, _lib.ptr(Sy), _lib.ptr(ws),
                                             ws.numel(), st), "rvo_ba_assemble")
            else:
                Sy.zero_()
            if n6:
                reduce_system(Sy, group)
            # E == 0 ranks still solve and retract their pose replicas (E=1 workspace, no patches touched)
            _lib.check(L.rvo_ba_solve_poses(_lib.ptr(pv), _lib.ptr(Sy), t0, t1, _lib.ptr(ws), ws.numel(), st)
                       if E == 0 else
                       L.rvo_ba_solve(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(Sy), E, qv.shape[0], P, t0, t1,
                                      _lib.ptr(ws), ws.numel(), st), "rvo_ba_solve")
    return []


def exchange_depths_owned(patches_, owners, frame_lo, frame_hi, rank, group=None):
    """exchange_depths for an explicit per-frame owner table (Ramp_vo keeps one: ownership follows a frame through
    keyframe removals, which renumber the frames).  patches_ [N, M, 3, P, P]; owners: a list of ranks or a DEVICE
    int tensor [N] (no host->device copy on the per-update path)."""
    if frame_hi <= frame_lo or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    sl = patches_[frame_lo:frame_hi, :, 2]                                   # [F, M, P, P] view
    if torch.is_tensor(owners):
        mine = (owners[frame_lo:frame_hi] == rank).to(sl.dtype).view(-1, 1, 1, 1)
    else:
        mine = torch.tensor([1.0 if o == rank else 0.0 for o in owners[frame_lo:frame_hi]], device=patches_.device,
                            dtype=sl.dtype).view(-1, 1, 1, 1)
    buf = (sl * mine).contiguous()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    sl.copy_(buf)


def sharded_BA_fused(poses, patches, intrinsics, coords, delta, weight, ht, wd, lmbda, ii, jj, kk, t0, t1,
                     iterations=2, weight_out=None, group=None, plan=None):
    """sharded_BA with the target formation / filter_features folded into the assembly (rvo_ba_assemble_fused):
    coords [1,E,2,P,P], delta / weight [1,E,2] of THIS rank's edges.  Collective per iteration: one all-reduce of
    [S | y] (6N x (6N+1) fp32)."""
    _lib.require_cuda(poses, patches, intrinsics, coords, delta, weight, lmbda, ii, jj, kk)
    L = _lib.lib()
    P = patches.shape[-1]
    pv, qv, kv = poses.view(-1, 7), patches.view(-1, 3, P, P), intrinsics.view(-1, 4)
    cv = coords.to(torch.float32).contiguous().view(-1, 2, P, P)
    dv = delta.to(torch.float32).contiguous().view(-1, 2)
    wv = weight.to(torch.float32).contiguous().view(-1, 2)
    lm = lmbda.to(torch.float32).contiguous().view(-1)
    ii, jj, kk = [t.to(torch.int64).contiguous() for t in (ii, jj, kk)]
    E, N = ii.numel(), t1 - t0
    n6 = 6 * N
    dev = pv.device
    ws = _lib.Workspace.get(dev, L.rvo_ba_ws_bytes(max(E, 1), qv.shape[0], N), "ba_sharded")
    Sy = torch.zeros(max(n6, 1), n6 + 1, dtype=torch.float32, device=dev)
    st = _lib.stream_ptr(dev)
    with torch.cuda.device(dev):
        if E and plan is not None:
            # the update operator already grouped these edges by (kk, jj): the workspace starts with a plan of the
            # same layout (rvo_plan_bytes), so the sort is replaced by one device copy
            nb = L.rvo_plan_bytes(E)
            ws[:nb].copy_(plan.view(torch.uint8).reshape(-1)[:nb])
        elif E:
            _lib.check(L.rvo_ba_plan(_lib.ptr(kk), _lib.ptr(jj), E, pv.shape[0], qv.shape[0], N,
                                     _lib.ptr(ws), ws.numel(), st), "rvo_ba_plan")
        for it in range(iterations):
            if E:
                _lib.check(L.rvo_ba_assemble_fused(
                    _lib.ptr(pv), _lib.ptr(qv), _lib.ptr(kv), _lib.ptr(cv), _lib.ptr(dv), _lib.ptr(wv), float(ht),
                    float(wd), _lib.ptr(weight_out) if (weight_out is not None and it == 0) else None, _lib.ptr(lm),
                    _lib.ptr(ii), _lib.ptr(jj), E, qv.shape[0], P, t0, t1, _lib.ptr(Sy), _lib.ptr(ws), ws.numel(),
                    st), "rvo_ba_assemble_fused")
            else:
                Sy.zero_()
            if n6:
                reduce_system(Sy, group)
            _lib.check(L.rvo_ba_solve_poses(_lib.ptr(pv), _lib.ptr(Sy), t0, t1, _lib.ptr(ws), ws.numel(), st)
                       if E == 0 else
                       L.rvo_ba_solve(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(Sy), E, qv.shape[0], P, t0, t1,
                                      _lib.ptr(ws), ws.numel(), st), "rvo_ba_solve")
    return []
