"""CPU: the numpy restatement checked against brute force, algebraic identities and itself."""
import numpy as np
import pytest

from oracle import ref_ops as O
from rampvo_b200 import synth
from tests.util import perturb_poses, targets_from_reprojection


def test_graph_sizes_match_reference_replay():
    # SURVEY.md section 8 table (replay of Ramp_vo.py:312-325,273)
    for cfg, nf, E, U, pairs in [("cfg1", 8, 2048, 256, 64), ("fast", 40, 13488, 768, 281),
                                 ("default", 40, 45312, 2112, 472)]:
        M, l, r, _ = synth.CONFIGS[cfg]
        ii, jj, kk = synth.replay_graph(M, l, r, nf)
        assert len(ii) == E and len(np.unique(kk)) == U
        assert len(np.unique(ii * 100000 + jj)) == pairs
        assert (ii == kk // M).all()


def test_neighbors_bruteforce():
    rng = np.random.default_rng(0)
    kk = rng.integers(0, 7, 200)
    jj = rng.integers(0, 5, 200)   # many ties -> stability matters
    ix, jx = O.neighbors(kk, jj)
    for k in np.unique(kk):
        idx = [e for e in range(200) if kk[e] == k]
        idx.sort(key=lambda e: jj[e])  # python sort is stable, like std::stable_sort (ba.cpp:85)
        for n, e in enumerate(idx):
            assert ix[e] == (idx[n - 1] if n > 0 else -1)
            assert jx[e] == (idx[n + 1] if n + 1 < len(idx) else -1)


def test_patchify_and_corr_small_bruteforce():
    rng = np.random.default_rng(1)
    net = rng.standard_normal((1, 4, 9, 11)).astype(np.float32)
    coords = np.array([[[2.25, 3.5], [0.1, 0.2], [10.7, 8.9], [-3.0, 4.0]]], np.float32)
    raw = O.patchify_raw(net, coords, 1)
    assert raw.shape == (1, 4, 4, 4, 4)
    assert raw[0, 0, 2, 1, 1] == net[0, 2, 3, 2]          # centre of the 4x4 window = floor
    assert (raw[0, 3, :, :, :2] == 0).all()              # x = -3 -> columns -4,-3 out of bounds
    bl = O.patchify(net, coords, 0)
    x, y = 2.25, 3.5
    exp = (0.5 * 0.75 * net[0, :, 3, 2] + 0.5 * 0.25 * net[0, :, 3, 3] + 0.5 * 0.75 * net[0, :, 4, 2] +
           0.5 * 0.25 * net[0, :, 4, 3])
    np.testing.assert_allclose(bl[0, 0, :, 0, 0], exp, rtol=1e-6)

    f1 = rng.standard_normal((3, 8, 3, 3))
    f2 = rng.standard_normal((2, 8, 12, 14))
    c = rng.uniform(-2, 15, (5, 2, 3, 3)).astype(np.float32)
    ii = np.array([0, 1, 2, 1, 0])
    jj = np.array([1, 0, 1, 1, 0])
    out = O.corr(f1, f2, c, ii, jj, 2)
    assert out.shape == (5, 5, 5, 3, 3)
    # brute force one entry: edge 3, pixel (1,2), x-offset 4, y-offset 0
    e, i0, j0, b, a = 3, 1, 2, 4, 0
    x, y = float(c[e, 0, i0, j0]), float(c[e, 1, i0, j0])
    fx, fy = int(np.floor(x)), int(np.floor(y))
    dx = float(np.float32(x) - np.floor(np.float32(x)))
    dy = float(np.float32(y) - np.floor(np.float32(y)))

    def dot(yy, xx):
        if 0 <= yy < 12 and 0 <= xx < 14:
            return float(f1[ii[e], :, i0, j0] @ f2[jj[e], :, yy, xx])
        return 0.0
    v = ((1 - dx) * (1 - dy) * dot(fy + a - 2, fx + b - 2) + dx * (1 - dy) * dot(fy + a - 2, fx + b - 1) +
         (1 - dx) * dy * dot(fy + a - 1, fx + b - 2) + dx * dy * dot(fy + a - 1, fx + b - 1))
    np.testing.assert_allclose(out[e, b, a, i0, j0], v, rtol=1e-9, atol=1e-12)


def test_transform_identities():
    prob = synth.make_problem("cfg1", 8, seed=3)
    ii, kk = prob["ii"], prob["kk"]
    c, d, v = O.transform(prob["poses"], prob["patches"], prob["intrinsics"], ii, ii, kk)
    # i -> i is the identity map on patch pixel coordinates
    np.testing.assert_allclose(c[..., 0], prob["patches"][kk, 0], atol=1e-4)
    np.testing.assert_allclose(c[..., 1], prob["patches"][kk, 1], atol=1e-4)
    # numerical Jacobian wrt the depth of the patch
    c0, _, (Ji, Jj, Jz) = O.transform(prob["poses"], prob["patches"], prob["intrinsics"], ii,
                                      prob["jj"], kk, jacobian=True)
    eps = 1e-6
    p2 = prob["patches"].astype(np.float64).copy()
    p2[:, 2] += eps
    c1 = O.transform(prob["poses"], p2, prob["intrinsics"], ii, prob["jj"], kk)[0]
    num = (c1[:, 1, 1] - c0[:, 1, 1]) / eps
    np.testing.assert_allclose(Jz[..., 0], num, rtol=1e-3, atol=1e-4)


def test_ba_jacobians_numerically():
    """Ji / Jj of the CUDA-BA restatement are derivatives of the projection wrt LEFT perturbations
    T <- Exp(xi) T of pose j / pose i (sign conventions of ba_cuda.cu:344-366)."""
    prob = synth.make_problem("cfg1", 8, seed=4)
    tgt = targets_from_reprojection(prob, O)
    sel = np.nonzero(prob["ii"] != prob["jj"])[0][:50]
    n = len(sel)
    # give every edge its own copy of pose i and pose j so they can be perturbed independently
    poses = np.concatenate([prob["poses"][prob["ii"][sel]], prob["poses"][prob["jj"][sel]]]).astype(np.float64)
    ii, jj, kk = np.arange(n), n + np.arange(n), prob["kk"][sel]
    args = (prob["patches"], prob["intrinsics"], tgt[sel], prob["weight"][sel], ii, jj, kk)
    r0, w, Ji, Jj, Jz = O.ba_linearise(poses, *args)
    eps = 1e-6
    for a in range(6):
        xi = np.zeros(6)
        xi[a] = eps
        for frames, J, sign in ((jj, Jj, 1.0), (ii, Ji, -1.0)):
            p = poses.copy()
            for f in frames:
                t, q = O.retr_se3(xi, p[f, :3], p[f, 3:])
                p[f, :3], p[f, 3:] = t, q
            r1 = O.ba_linearise(p, *args)[0]
            dproj = -(r1 - r0) / eps            # r = target - proj
            np.testing.assert_allclose(sign * J[:, :, a], dproj, rtol=2e-3, atol=2e-3)


def test_ba_reduces_residual_and_recovers_poses():
    prob = synth.make_problem("cfg1", 8, seed=5, noise_px=0.0)
    tgt = targets_from_reprojection(prob, O)
    w = np.ones_like(prob["weight"])
    noisy = perturb_poses(prob)

    def cost(poses, patches):
        r, wm, *_ = O.ba_linearise(poses, patches, prob["intrinsics"], tgt, w, prob["ii"], prob["jj"], prob["kk"])
        return float((wm * r * r).sum())
    c0 = cost(noisy, prob["patches"])
    p, q = noisy, prob["patches"]
    for _ in range(4):
        p, q = O.ba(p, q, prob["intrinsics"], tgt, w, 1e-4, prob["ii"], prob["jj"], prob["kk"],
                    prob["t0"], prob["t1"], iterations=2)
    c1 = cost(p, q)
    assert c1 < 1e-3 * c0
    # structure-only branch (t1 == t0) also decreases the cost
    p2, q2 = O.ba(noisy, prob["patches"], prob["intrinsics"], tgt, w, 1e-4, prob["ii"], prob["jj"],
                  prob["kk"], 3, 3, iterations=2)
    assert (p2 == noisy.astype(np.float64)).all()
    assert cost(p2, q2) < c0


def test_ba_schur_matches_full_system_solve():
    """Solving the full [[B,E],[E^T,C+lambda]] system gives the same dX as the Schur path."""
    prob = synth.make_problem("cfg1", 8, seed=6)
    tgt = targets_from_reprojection(prob, O)
    s = O.ba_system(perturb_poses(prob), prob["patches"], prob["intrinsics"], tgt, prob["weight"],
                    1e-4, prob["ii"], prob["jj"], prob["kk"], prob["t0"], prob["t1"])
    n6, M = s["B"].shape[0], len(s["C"])
    H = np.block([[s["B"], s["E"]], [s["E"].T, np.diag(s["C"] + 1e-4)]])
    g = np.concatenate([s["v"], s["u"]])
    full = np.linalg.solve(H, g)
    dX = np.linalg.solve(s["S"], s["y"])
    np.testing.assert_allclose(dX, full[:n6], rtol=1e-6, atol=1e-9)
    dZ = s["Q"] * (s["u"] - s["E"].T @ dX)
    np.testing.assert_allclose(dZ, full[n6:], rtol=1e-6, atol=1e-9)
