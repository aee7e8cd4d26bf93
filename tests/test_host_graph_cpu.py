"""CPU: the host-side patch-graph bookkeeping of rampvo_b200.Ramp_vo (append_factors / remove_factors /
edge rules / pair counts / hidden-state ping-pong buffers) against an independent replay of the
reference's rules (ramp/Ramp_vo.py:194-208,312-325,273 via rampvo_b200.synth.replay_graph).  No kernel
is launched: these paths are plain tensor plumbing and run on CPU tensors."""
import numpy as np
import pytest
import torch

from rampvo_b200 import synth
from rampvo_b200.Ramp_vo import Ramp_vo
from rampvo_b200.config import preset
from rampvo_b200.net import VONet


@pytest.fixture(scope="module")
def vo_cpu():
    torch.manual_seed(0)
    train_cfg = {"event_bias": True, "input_mode": "MultiScale", "num_event_bins": 5}
    cfg = preset("cfg1")
    cfg.BUFFER_SIZE = 64
    return Ramp_vo(cfg, VONet(train_cfg), train_cfg, ht=64, wd=64, device="cpu", use_graphs=False)


def test_edge_rules_match_the_reference_replay(vo_cpu):
    vo = vo_cpu
    M, life, removal = vo.M, vo.cfg.PATCH_LIFETIME, vo.cfg.REMOVAL_WINDOW
    assert (M, life, removal) == synth.CONFIGS["cfg1"][:3]
    tag = lambda kk, jj: (kk * 1000 + jj).float()
    for n in range(1, 31):
        vo.n, vo.m = n, n * M
        E0 = vo.ii.numel()
        vo.append_factors(*vo._edges_forw())
        vo.append_factors(*vo._edges_back())
        # new edges start from a zero hidden state (Ramp_vo.py:199-200); tag every row with its edge
        assert (vo.net[0, E0:] == 0).all()
        vo.net[0, E0:, 0] = tag(vo.kk[E0:], vo.jj[E0:])
        lim = n - removal
        vo.remove_factors(vo.ii < lim, lambda i, j: i < lim)
        ii, jj, kk = synth.replay_graph(M, life, removal, n)
        assert np.array_equal(vo.ii.numpy(), ii) and np.array_equal(vo.jj.numpy(), jj)
        assert np.array_equal(vo.kk.numpy(), kk)
        # the hidden state follows its edge through the compaction into the other ping-pong buffer
        assert vo.net.shape == (1, len(ii), vo.DIM)
        assert torch.equal(vo.net[0, :, 0], tag(vo.kk, vo.jj))
        # host-side pair counts (what lets removals avoid a device->host sync) stay exact
        key = ii * 10000 + jj
        u, c = np.unique(key, return_counts=True)
        want = {(int(k) // 10000, int(k) % 10000): int(v) for k, v in zip(u, c)}
        assert vo._pair_counts() == want
    assert vo.ii.numel() == synth.make_problem("cfg1", 30)["E"]


def test_remove_factors_with_a_device_mask_only(vo_cpu):
    """the reference signature (boolean mask, no predicate) still works and recounts the pairs lazily"""
    vo = vo_cpu
    E = vo.ii.numel()
    jmax = int(vo.jj.max())
    m = vo.jj == jmax
    gone = int(m.sum())
    assert 0 < gone < E
    vo.remove_factors(m)
    assert vo.ii.numel() == E - gone and not (vo.jj == jmax).any()
    assert vo.net.shape[1] == E - gone and sum(vo._pair_counts().values()) == E - gone


def test_keyframe_drop_matches_a_restatement_of_the_reference(vo_cpu):
    """Ramp_vo._keyframe_finish with a forced drop against a numpy restatement of ramp/Ramp_vo.py:243-274
    (frame-by-frame shift loop, boolean-indexed renumbering, removal window)."""
    vo = vo_cpu
    M, life, removal, KI = vo.M, vo.cfg.PATCH_LIFETIME, vo.cfg.REMOVAL_WINDOW, vo.cfg.KEYFRAME_INDEX
    n = 30
    ii, jj, kk = (torch.from_numpy(a) for a in synth.replay_graph(M, life, removal, n))
    vo.n, vo.m = n, n * M
    vo.ii, vo.jj, vo.kk = ii.clone(), jj.clone(), kk.clone()
    vo._pair_cnt = None
    vo._net_bufs, vo._net_cur = [None, None], 0
    vo.net = torch.zeros(1, 0, vo.DIM)
    vo._net_reserve(ii.numel())
    vo.net = vo._net_bufs[0][:, :ii.numel()]
    vo.net[0, :, 0] = (kk * 1000 + jj).float()
    g = torch.Generator().manual_seed(3)
    for buf in (vo.poses_, vo.patches_, vo.intrinsics_, vo.imap_, vo._gmap_store, vo._fmap1_store, vo._fmap2_store):
        buf.copy_(torch.rand(buf.shape, generator=g).to(buf.dtype))
    vo.poses_[:, 3:] = torch.nn.functional.normalize(vo.poses_[:, 3:] + 0.1, dim=-1)
    vo.tstamps_[:n] = torch.arange(n)
    vo.colors_.copy_(torch.randint(0, 255, vo.colors_.shape, generator=g).to(torch.uint8))
    vo.delta = {}
    snap = {k: getattr(vo, k).clone() for k in ("tstamps_", "colors_", "poses_", "patches_", "intrinsics_", "imap_",
                                                "_gmap_store", "_fmap1_store", "_fmap2_store")}
    old_thresh = vo.cfg.KEYFRAME_THRESH
    vo.cfg.KEYFRAME_THRESH = 1e9
    try:
        vo._keyframe_finish([0.0, 1.0, 0.0, 1.0])
    finally:
        vo.cfg.KEYFRAME_THRESH = old_thresh

    # ---- restatement
    k = n - KI
    keep = ~((ii == k) | (jj == k))
    ri, rj, rk = ii[keep].clone(), jj[keep].clone(), kk[keep].clone()
    rk[ri > k] -= M
    ri[ri > k] -= 1
    rj[rj > k] -= 1
    mem = vo.mem
    exp = {name: t.clone() for name, t in snap.items()}
    gview = exp["_gmap_store"].view(mem, M, vo.P, vo.P, 128)
    for i in range(k, n - 1):
        for name in ("tstamps_", "colors_", "poses_", "patches_", "intrinsics_"):
            exp[name][i] = exp[name][i + 1]
        exp["imap_"][i % mem] = exp["imap_"][(i + 1) % mem]
        gview[i % mem] = gview[(i + 1) % mem]
        exp["_fmap1_store"][i % mem] = exp["_fmap1_store"][(i + 1) % mem]
        exp["_fmap2_store"][i % mem] = exp["_fmap2_store"][(i + 1) % mem]
    n2 = n - 1
    keep2 = ~(ri < n2 - removal)                     # ix[kk] == ii for the identity index map
    ri, rj, rk = ri[keep2], rj[keep2], rk[keep2]

    assert vo.n == n2 and vo.m == n2 * M
    assert torch.equal(vo.ii, ri) and torch.equal(vo.jj, rj) and torch.equal(vo.kk, rk)
    for name, t in exp.items():
        assert torch.equal(getattr(vo, name), t), name
    assert list(vo.delta) == [k] and vo.delta[k][0] == k - 1
    # hidden states followed their edges (tags carry the OLD patch / frame numbers)
    old_tag = (kk[keep][keep2] * 1000 + jj[keep][keep2]).float()
    assert torch.equal(vo.net[0, :, 0], old_tag)
    assert sum(vo._pair_counts().values()) == ri.numel()


def test_deferred_keyframe_drop_flushes_to_the_same_graph(vo_cpu):
    """pipeline mode leaves a keyframe drop to the fused patch-graph step of the next frame (rvo_edges_step drop_k):
    _keyframe_finish(defer_removal=True) only records it (host pair counts already updated); if no new frame comes
    (sync(), terminate()) the tensor-op path applies it — the graph must equal the one of an immediate drop."""
    vo = vo_cpu
    M, life, removal, KI = vo.M, vo.cfg.PATCH_LIFETIME, vo.cfg.REMOVAL_WINDOW, vo.cfg.KEYFRAME_INDEX
    n = 30
    ii, jj, kk = (torch.from_numpy(a) for a in synth.replay_graph(M, life, removal, n))

    def reset():
        vo.n, vo.m = n, n * M
        vo.ii, vo.jj, vo.kk = ii.clone(), jj.clone(), kk.clone()
        vo._pair_cnt, vo._min_src = None, 0
        vo._net_bufs, vo._net_cur = [None, None], 0
        vo.net = torch.zeros(1, 0, vo.DIM)
        vo._net_reserve(ii.numel())
        vo.net = vo._net_bufs[0][:, :ii.numel()]
        vo.net[0, :, 0] = (kk * 1000 + jj).float()
        vo.tstamps_[:n] = torch.arange(n)
        vo.delta = {}

    old_thresh = vo.cfg.KEYFRAME_THRESH
    vo.cfg.KEYFRAME_THRESH = 1e9
    try:
        reset()
        vo._keyframe_finish([0.0, 1.0, 0.0, 1.0])
        want = (vo.ii.clone(), vo.jj.clone(), vo.kk.clone(), vo.net[0, :, 0].clone(), dict(vo._pair_counts()), vo.n)
        reset()
        vo._keyframe_finish([0.0, 1.0, 0.0, 1.0], defer_removal=True)
        k = n - KI
        assert vo._pending_drop == k and vo._pending_lim == vo.n - removal
        assert torch.equal(vo.ii, ii) and torch.equal(vo.jj, jj)        # the device lists wait for the fused step
        assert sum(vo._pair_cnt.values()) == int((~((ii == k) | (jj == k))).sum())   # the host counts do not
        vo.sync()
    finally:
        vo.cfg.KEYFRAME_THRESH = old_thresh
    assert vo._pending_drop == -1 and vo._pending_lim is None and vo.n == want[5]
    assert torch.equal(vo.ii, want[0]) and torch.equal(vo.jj, want[1]) and torch.equal(vo.kk, want[2])
    assert torch.equal(vo.net[0, :, 0], want[3])
    assert vo._pair_counts() == want[4]


def _q_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def _q_rot(q, v):
    uv = 2 * np.cross(q[:3], v)
    return v + q[3] * uv + np.cross(q[:3], uv)


def _se3_mul(A, B):          # [t, q xyzw]: (A * B)(x) = A(B(x))
    return np.concatenate([A[:3] + _q_rot(A[3:], B[:3]), _q_mul(A[3:], B[3:])])


def _se3_inv(A):
    qi = A[3:] * np.array([-1, -1, -1, 1.0])
    return np.concatenate([-_q_rot(qi, A[:3]), qi])


def test_terminate_interpolates_dropped_frames_like_the_reference(vo_cpu):
    """terminate() / get_pose (ramp/Ramp_vo.py:155-173): poses of dropped frames are rebuilt from the stored
    relative motions dP = T_k * T_{k-1}^-1 and everything is returned as camera-to-world."""
    vo = vo_cpu
    rng = np.random.default_rng(5)
    kept = [0, 1, 2, 4, 5, 8]                      # timestamps that still own a keyframe slot
    P = rng.normal(0, 0.3, (len(kept), 7))
    P[:, 3:] /= np.linalg.norm(P[:, 3:], axis=1, keepdims=True)
    vo.n, vo.counter = len(kept), 9
    vo.sync()
    vo.poses_[:vo.n] = torch.from_numpy(P).float()
    vo.tstamps_[:vo.n] = torch.tensor(kept)
    vo.tlist = list(range(100, 109))
    D = {}
    for t, t0 in ((3, 2), (6, 5), (7, 6)):          # 7 chains through the dropped 6
        d = rng.normal(0, 0.1, 7)
        d[3:] = d[3:] * 0.1 + np.array([0, 0, 0, 1.0])
        d[3:] /= np.linalg.norm(d[3:])
        D[t] = (t0, d)
    from rampvo_b200.lietorch import SE3
    vo.delta = {t: (t0, SE3(torch.from_numpy(d).float()[None])[0]) for t, (t0, d) in D.items()}
    poses, tstamps = vo.terminate()

    traj = {t: P[i] for i, t in enumerate(kept)}

    def get(t):
        if t in traj:
            return traj[t]
        t0, d = D[t]
        return _se3_mul(d, get(t0))
    exp = np.stack([_se3_inv(get(t)) for t in range(9)])
    assert poses.shape == (9, 7) and np.array_equal(tstamps, np.arange(100, 109, dtype=float))
    # quaternion sign is free
    sgn = np.sign((poses[:, 3:] * exp[:, 3:]).sum(1, keepdims=True))
    assert np.abs(poses[:, :3] - exp[:, :3]).max() < 1e-5
    assert np.abs(poses[:, 3:] * sgn - exp[:, 3:]).max() < 1e-5


def test_tile_layout_weight_permutation_is_the_same_linear_map():
    """The tcgen05 corr kernel writes [E, 1008] in the tile layout; the update operator reads it through
    column-permuted first-layer weights (net.py: corr0t).  On CPU: the permuted layer applied to the tile
    layout equals the reference layer (ramp/net.py:52, Linear(882, 384)) applied to the reference layout."""
    from rampvo_b200 import altcorr
    from rampvo_b200.net import Update
    torch.manual_seed(1)
    up = Update(3)
    W = up._fused_weights()
    cols, ref = altcorr.tile_layout_index(2)
    assert len(cols) == 882 and sorted(ref.tolist()) == list(range(882)) and int(cols.max()) < 1008
    c_ref = torch.randn(50, 882).half()
    tile = torch.zeros(50, 1008, dtype=torch.float16)
    tile[:, cols] = c_ref[:, ref]
    a = c_ref.double() @ W["corr0"][0].double().t()
    b = tile.double() @ W["corr0t"].double().t()
    assert torch.allclose(a, b, rtol=0, atol=1e-9)
    # pad columns of the tile layout carry zero weights
    pad = torch.ones(1008, dtype=torch.bool)
    pad[cols] = False
    assert (W["corr0t"][:, pad] == 0).all()
