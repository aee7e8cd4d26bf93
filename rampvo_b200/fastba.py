"""Drop-in for ramp.fastba (ramp/fastba/ba.py:4-8, cuda_ba ramp/fastba/ba.cpp:183-188)."""
import torch

from . import _lib


def _i64(t):
    return t.to(torch.int64).contiguous()


def neighbors(ii, jj, kmax=0, jmax=0):
    """cuda_ba.neighbors(ii, jj) (ba.cpp:59-97; called as neighbors(kk, jj) at net.py:77): for each
    edge the previous / next edge with the same first key, ordered by the second key (stable);
    -1 at the ends.  Runs on the device — no host round trip.  kmax/jmax (optional, extension):
    exclusive upper bounds of the two keys, which shorten the radix sort."""
    _lib.require_cuda(ii, jj)
    E = ii.numel()
    ii, jj = _i64(ii), _i64(jj)
    ix = torch.empty(E, dtype=torch.int64, device=ii.device)
    jx = torch.empty(E, dtype=torch.int64, device=ii.device)
    if E == 0:
        return [ix, jx]
    L = _lib.lib()
    nb = L.rvo_neighbors_ws_bytes(E)
    ws = _lib.Workspace.get(ii.device, nb, "neighbors")
    with torch.cuda.device(ii.device):
        _lib.check(L.rvo_neighbors(_lib.ptr(ii), _lib.ptr(jj), E, kmax, jmax, _lib.ptr(ix),
                                   _lib.ptr(jx), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                   "rvo_neighbors")
    return [ix, jx]


def _flat(t, name, last):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError("fastba: %s must be a contiguous float32 tensor" % name)
    return t.view(-1, *last)


def BA(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, M, iterations,
       eff_impl=False, plan=None, t0_dev=None):
    """ramp.fastba.BA (ba.py:7 -> cuda_ba.forward, ba_cuda.cu:433-582).  Mutates `poses`
    (rows t0..t1-1) and the inverse depths in `patches` IN PLACE and returns [] like the reference.
    `poses` may be a lietorch SE3 (its .data is used, ba.py:8).  `plan` (extension): the device
    buffer of an rvo_graph_plan(kk, jj) built for the same edge list — skips the edge sort;
    `t0_dev` (extension, needs `plan`): int32 device scalar holding t0, read by the kernels instead of
    the host value so that a captured CUDA graph stays valid while the window slides (t1 - t0 fixed)."""
    poses = getattr(poses, "data", poses)
    _lib.require_cuda(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk)
    P = patches.shape[-1]
    pv = _flat(poses, "poses", (7,))
    qv = _flat(patches, "patches", (3, P, P))
    kv = _flat(intrinsics, "intrinsics", (4,))
    tv = target.to(torch.float32).contiguous().view(-1, 2)
    wv = weight.to(torch.float32).contiguous().view(-1, 2)
    lm = lmbda.to(torch.float32).contiguous().view(-1)
    E = ii.numel()
    if tv.shape[0] != E or wv.shape[0] != E or jj.numel() != E or kk.numel() != E:
        raise RuntimeError("fastba.BA: target/weight/ii/jj/kk disagree on the number of edges")
    ii, jj, kk = _i64(ii), _i64(jj), _i64(kk)
    L = _lib.lib()
    n_poses, n_patches = pv.shape[0], qv.shape[0]
    nb = L.rvo_ba_ws_bytes(E, n_patches, max(t1 - t0, 0))
    ws = _lib.Workspace.get(pv.device, nb, "ba")
    if plan is not None and t0_dev is not None:
        with torch.cuda.device(pv.device):
            _lib.check(L.rvo_ba_forward_dyn(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(kv), _lib.ptr(tv),
                                            _lib.ptr(wv), _lib.ptr(lm), _lib.ptr(ii), _lib.ptr(jj),
                                            _lib.ptr(plan), E, n_poses, n_patches, P, int(t1 - t0),
                                            _lib.ptr(t0_dev), int(iterations), _lib.ptr(ws), ws.numel(),
                                            _lib.stream_ptr()), "rvo_ba_forward_dyn")
        return []
    if plan is not None:
        with torch.cuda.device(pv.device):
            _lib.check(L.rvo_ba_forward_planned(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(kv), _lib.ptr(tv),
                                                _lib.ptr(wv), _lib.ptr(lm), _lib.ptr(ii), _lib.ptr(jj),
                                                _lib.ptr(plan), E, n_poses, n_patches, P, int(t0), int(t1),
                                                int(iterations), _lib.ptr(ws), ws.numel(),
                                                _lib.stream_ptr()), "rvo_ba_forward_planned")
        return []
    with torch.cuda.device(pv.device):
        _lib.check(L.rvo_ba_forward(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(kv), _lib.ptr(tv),
                                    _lib.ptr(wv), _lib.ptr(lm), _lib.ptr(ii), _lib.ptr(jj),
                                    _lib.ptr(kk), E, n_poses, n_patches, P, int(M), int(t0), int(t1),
                                    int(iterations), int(bool(eff_impl)), _lib.ptr(ws), ws.numel(),
                                    _lib.stream_ptr()), "rvo_ba_forward")
    return []


def BA_fused(poses, patches, intrinsics, coords, delta, weight, ht, wd, lmbda, ii, jj, plan, n_free, t0_dev,
             iterations, weight_out=None):
    """fastba.BA with the caller-side target formation of ramp/Ramp_vo.py:288-296 folded into the kernels' edge
    load: target = coords[..., P//2, P//2] + delta, weight zeroed where the target leaves [0,wd] x [0,ht]
    (ramp/utils.py:557-570).  coords [1,E,2,P,P] fp32; delta / weight [1,E,2] fp32; plan: rvo_graph_plan(kk, jj);
    the window is [t0_dev[0], t0_dev[0] + n_free).  weight_out [1,E,2] receives the filtered confidences."""
    poses = getattr(poses, "data", poses)
    _lib.require_cuda(poses, patches, intrinsics, coords, delta, weight, lmbda, ii, jj, plan, t0_dev)
    P = patches.shape[-1]
    pv = _flat(poses, "poses", (7,))
    qv = _flat(patches, "patches", (3, P, P))
    kv = _flat(intrinsics, "intrinsics", (4,))
    E = ii.numel()
    cv = _flat(coords, "coords", (2, P, P))
    dv = _flat(delta, "delta", (2,))
    wv = _flat(weight, "weight", (2,))
    if cv.shape[0] != E or dv.shape[0] != E or wv.shape[0] != E or jj.numel() != E:
        raise RuntimeError("fastba.BA_fused: coords/delta/weight/ii/jj disagree on the number of edges")
    if weight_out is not None:
        weight_out = _flat(weight_out, "weight_out", (2,))
    lm = lmbda.to(torch.float32).contiguous().view(-1)
    ii, jj = _i64(ii), _i64(jj)
    L = _lib.lib()
    nb = L.rvo_ba_ws_bytes(E, qv.shape[0], max(int(n_free), 0))
    ws = _lib.Workspace.get(pv.device, nb, "ba")
    with torch.cuda.device(pv.device):
        _lib.check(L.rvo_ba_forward_fused(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(kv), _lib.ptr(cv), _lib.ptr(dv),
                                          _lib.ptr(wv), float(ht), float(wd), _lib.ptr(weight_out), _lib.ptr(lm),
                                          _lib.ptr(ii), _lib.ptr(jj), _lib.ptr(plan), E, qv.shape[0], P,
                                          int(n_free), _lib.ptr(t0_dev), int(iterations), _lib.ptr(ws), ws.numel(),
                                          _lib.stream_ptr()), "rvo_ba_forward_fused")
    return []


def reproject(poses, patches, intrinsics, ii, jj, kk):
    """cuda_ba.reproject (ba.cpp:49-57, ba_cuda.cu:379-429,585-617): coords [1,E,2,P,P]; no depth
    clamp and intrinsics[0] for every frame, exactly like the reference kernel."""
    poses = getattr(poses, "data", poses)
    _lib.require_cuda(poses, patches, intrinsics, ii, jj, kk)
    P = patches.shape[-1]
    pv = _flat(poses, "poses", (7,))
    qv = _flat(patches, "patches", (3, P, P))
    kv = _flat(intrinsics, "intrinsics", (4,))
    E = ii.numel()
    ii, jj, kk = _i64(ii), _i64(jj), _i64(kk)
    out = torch.empty(1, E, 2, P, P, dtype=torch.float32, device=pv.device)
    with torch.cuda.device(pv.device):
        _lib.check(_lib.lib().rvo_reproject(_lib.ptr(pv), _lib.ptr(qv), _lib.ptr(kv), _lib.ptr(ii),
                                            _lib.ptr(jj), _lib.ptr(kk), E, P, _lib.ptr(out),
                                            _lib.stream_ptr()), "rvo_reproject")
    return out
