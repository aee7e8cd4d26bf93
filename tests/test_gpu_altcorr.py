"""GPU parity: altcorr (patchify / corr) through the C-ABI vs the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import ref_ops as O
from rampvo_b200 import altcorr, synth

pytestmark = pytest.mark.gpu


def _coords(rng, M, W, H, frac=True):
    x = rng.uniform(-3, W + 3, M)
    y = rng.uniform(-3, H + 3, M)
    if not frac:
        x, y = np.floor(x), np.floor(y)
    return np.stack([x, y], -1).astype(np.float32)[None]


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("radius", [0, 1])
def test_patchify_raw_bit_exact(dtype, channels_last, radius):
    rng = np.random.default_rng(0)
    net = rng.standard_normal((1, 12, 30, 40)).astype(dtype)
    coords = _coords(rng, 50, 40, 30)
    t = torch.from_numpy(net).cuda()
    if channels_last:
        t = t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    got = altcorr.patchify(t, torch.from_numpy(coords).cuda(), radius, mode="nearest").cpu().numpy()
    exp = O.patchify_raw(net, coords, radius)
    assert got.dtype == exp.dtype and (got == exp).all()       # bit-exact gather


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("radius", [0, 1])
def test_patchify_bilinear_bit_exact(dtype, radius):
    rng = np.random.default_rng(1)
    net = rng.standard_normal((1, 16, 30, 40)).astype(dtype)
    coords = _coords(rng, 64, 40, 30)
    got = altcorr.patchify(torch.from_numpy(net).cuda(), torch.from_numpy(coords).cuda(), radius)
    exp = O.patchify(net, coords, radius)
    assert got.dtype == torch.float32                           # correlation.py:57-66 promotes to fp32
    assert (got.cpu().numpy() == exp).all()                     # same association, no FMA => bit-exact
    # strided fp16 destination (the channels-last gmap ring): rounded once from the fp32 blend
    ring = torch.zeros(1, 64, 2 * radius + 1, 2 * radius + 1, 16, dtype=torch.float16, device="cuda")
    view = ring.permute(0, 1, 4, 2, 3)
    altcorr.patchify(torch.from_numpy(net).cuda(), torch.from_numpy(coords).cuda(), radius, out=view)
    assert (view.cpu().numpy() == exp.astype(np.float16)).all()


def test_patchify_hot_path_shapes():
    """The four calls of Patchifier.forward (net.py:190-202) at 120x160."""
    rng = np.random.default_rng(2)
    xs = rng.integers(1, 159, 96).astype(np.float32)
    ys = rng.integers(1, 119, 96).astype(np.float32)
    coords = np.stack([xs + ys / 120.0, ys], -1)[None].astype(np.float32)   # utils.py:212 quirk
    c = torch.from_numpy(coords).cuda()
    for C, R, dt in [(128, 1, np.float16), (384, 0, np.float16), (3, 1, np.float32)]:
        net = rng.standard_normal((1, C, 120, 160)).astype(dt)
        got = altcorr.patchify(torch.from_numpy(net).cuda(), c, R).cpu().numpy()
        assert (got == O.patchify(net, coords, R)).all()


def _corr_case(rng, E, Np, Nf, C, H, W, P, spread=1.0):
    ii = rng.integers(0, Np, E)
    jj = rng.integers(0, Nf, E)
    cx = rng.uniform(-6, W + 6, E)
    cy = rng.uniform(-6, H + 6, E)
    g = (np.arange(P) - P // 2) * spread
    coords = np.zeros((E, 2, P, P), np.float32)
    coords[:, 0] = cx[:, None, None] + g[None, None, :] + rng.normal(0, 0.05, (E, P, P))
    coords[:, 1] = cy[:, None, None] + g[None, :, None] + rng.normal(0, 0.05, (E, P, P))
    return ii, jj, coords


@pytest.mark.parametrize("radius", [1, 3])
def test_corr_fp32_nchw_matches_oracle(radius):
    """Reference layout (NCHW, fp32): generic kernel, fp32 accumulate; tolerance 1e-5 of the output scale."""
    rng = np.random.default_rng(3)
    C, H, W, P = 32, 24, 30, 3
    f1 = (rng.standard_normal((1, 20, C, P, P)) / np.sqrt(C)).astype(np.float32)
    f2 = (rng.standard_normal((1, 4, C, H, W)) / np.sqrt(C)).astype(np.float32)
    ii, jj, coords = _corr_case(rng, 300, 20, 4, C, H, W, P)
    got = altcorr.corr(torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda(),
                       torch.from_numpy(coords).cuda()[None], torch.from_numpy(ii).cuda(),
                       torch.from_numpy(jj).cuda(), radius)
    d = 2 * radius + 1
    assert got.shape == (1, 300, d, d, P, P)
    exp = O.corr(f1[0], f2[0], coords, ii, jj, radius)
    assert np.abs(got[0].cpu().numpy() - exp).max() < 1e-5 * max(1.0, np.abs(exp).max())


@pytest.mark.parametrize("spread", [1.0, 3.5])
@pytest.mark.parametrize("nlevels", [1, 2])
def test_corr_pyramid_fp16_tensor_core_path(spread, nlevels):
    """Ramp_vo.corr layout: channels-last fp16 ring buffers, radius 3, [E,882].  fp32 accumulate +
    one fp16 rounding: |err| <= 2^-10 * |value| + small absolute term.  spread=3.5 forces the
    per-pixel-window fallback (patch pixels further apart than the shared 10x10 window)."""
    rng = np.random.default_rng(4)
    Np, Nf, C, H, W, P, E = 40, 6, 128, 30, 40, 3, 500
    gmap, pyr = synth.make_features(Nf, Np, C, H, W, P, seed=4)
    pyr = pyr[:nlevels] if nlevels == 2 else pyr[:1]
    kk = rng.integers(0, 3 * Np, E)
    jj = rng.integers(0, 3 * Nf, E)
    _, _, coords = _corr_case(rng, E, Np, Nf, C, H, W, P, spread)
    g_t = torch.from_numpy(gmap).cuda().permute(0, 3, 1, 2)[None]             # [1,Np,C,P,P] view
    p_t = [torch.from_numpy(p).cuda().permute(0, 3, 1, 2)[None] for p in pyr]  # [1,Nf,C,H,W] views
    got = altcorr.corr_pyramid(g_t, p_t, torch.from_numpy(coords).cuda()[None],
                               torch.from_numpy(kk).cuda(), torch.from_numpy(jj).cuda(), Np, Nf, 3)
    assert got.shape == (1, E, 49 * 9 * nlevels) and got.dtype == torch.float16
    exp = O.corr_pyramid(gmap.transpose(0, 3, 1, 2), [p.transpose(0, 3, 1, 2) for p in pyr], coords,
                         kk, jj, Np, Nf, 3, scales=(1.0, 0.25)[:nlevels])
    g = got[0].float().cpu().numpy().astype(np.float64)
    assert (np.abs(g - exp) <= 2.0 ** -10 * np.abs(exp) + 2e-4).all()
    # the generic kernel (NCHW copies of the same data) agrees with the tensor-core path
    g2 = altcorr.corr_pyramid(g_t.contiguous(), [p.contiguous() for p in p_t],
                              torch.from_numpy(coords).cuda()[None], torch.from_numpy(kk).cuda(),
                              torch.from_numpy(jj).cuda(), Np, Nf, 3)
    assert (g2[0].float() - got[0].float()).abs().max().item() < 2e-3


def test_corr_matches_two_single_level_calls():
    """corr_pyramid == stack([corr(level1), corr(level2)], -1) as Ramp_vo.corr builds it (Ramp_vo.py:180-182)."""
    rng = np.random.default_rng(5)
    Np, Nf, C, H, W, P, E = 16, 4, 128, 32, 48, 3, 200
    gmap, pyr = synth.make_features(Nf, Np, C, H, W, P, seed=5)
    kk = rng.integers(0, Np, E)
    jj = rng.integers(0, Nf, E)
    _, _, coords = _corr_case(rng, E, Np, Nf, C, H, W, P)
    g_t = torch.from_numpy(gmap).cuda().permute(0, 3, 1, 2)[None]
    p_t = [torch.from_numpy(p).cuda().permute(0, 3, 1, 2)[None] for p in pyr]
    c_t = torch.from_numpy(coords).cuda()[None]
    k_t, j_t = torch.from_numpy(kk).cuda(), torch.from_numpy(jj).cuda()
    fused = altcorr.corr_pyramid(g_t, p_t, c_t, k_t, j_t, 0, 0, 3)
    c1 = altcorr.corr(g_t, p_t[0], c_t, k_t, j_t, 3)
    c2 = altcorr.corr(g_t, p_t[1], c_t / 4, k_t, j_t, 3)
    ref = torch.stack([c1, c2], -1).view(1, E, -1)
    assert (fused == ref).all()


def test_corr_edge_cases():
    g = torch.zeros(1, 4, 128, 3, 3, dtype=torch.float16, device="cuda")
    f = torch.zeros(1, 2, 128, 8, 8, dtype=torch.float16, device="cuda")
    empty = altcorr.corr(g, f, torch.zeros(1, 0, 2, 3, 3, device="cuda"),
                         torch.zeros(0, dtype=torch.long, device="cuda"),
                         torch.zeros(0, dtype=torch.long, device="cuda"), 3)
    assert empty.shape == (1, 0, 7, 7, 3, 3)
    # everything out of bounds / NaN coordinates -> zeros, no crash
    c = torch.full((1, 3, 2, 3, 3), 1e9, device="cuda")
    c[0, 1] = float("nan")
    c[0, 2] = -1e9
    z = altcorr.corr(g + 1, f + 1, c, torch.tensor([0, 1, 2], device="cuda"),
                     torch.tensor([0, 1, 0], device="cuda"), 3)
    assert torch.isfinite(z[0, [0, 2]]).all() and (z[0, [0, 2]] == 0).all()
    with pytest.raises(RuntimeError):
        altcorr.corr(g, f.float(), c, torch.tensor([0, 1, 2], device="cuda"),
                     torch.tensor([0, 1, 0], device="cuda"), 3)


def test_corr_full_size_properties():
    """default.yaml steady state (E = 45 312): linearity in the patch features and agreement of a
    random sample of edges with the oracle."""
    prob = synth.make_problem("default", 40, seed=0)
    E = prob["E"]
    gmap, pyr = synth.make_features(32, 96 * 32, seed=0)
    c = O.reproject(prob["poses"], prob["patches"], prob["intrinsics"], prob["ii"], prob["jj"],
                    prob["kk"]).astype(np.float32)
    g_t = torch.from_numpy(gmap).cuda().permute(0, 3, 1, 2)[None]
    p_t = [torch.from_numpy(p).cuda().permute(0, 3, 1, 2)[None] for p in pyr]
    c_t = torch.from_numpy(c).cuda()[None]
    k_t, j_t = torch.from_numpy(prob["kk"]).cuda(), torch.from_numpy(prob["jj"]).cuda()
    out = altcorr.corr_pyramid(g_t, p_t, c_t, k_t, j_t, 96 * 32, 32, 3)
    assert out.shape == (1, E, 882)
    out2 = altcorr.corr_pyramid(g_t * 2, p_t, c_t, k_t, j_t, 96 * 32, 32, 3)
    assert (out2.float() - 2 * out.float()).abs().max().item() < 2e-3          # linearity
    sel = np.random.default_rng(0).choice(E, 64, replace=False)
    exp = O.corr_pyramid(gmap.transpose(0, 3, 1, 2), [p.transpose(0, 3, 1, 2) for p in pyr], c[sel],
                         prob["kk"][sel], prob["jj"][sel], 96 * 32, 32, 3)
    got = out[0, torch.from_numpy(sel).cuda()].float().cpu().numpy()
    assert (np.abs(got - exp) <= 2.0 ** -10 * np.abs(exp) + 2e-4).all()


@pytest.mark.parametrize("spread", [1.0, 2.2])
def test_corr_tiles_tcgen05_matches_oracle(spread):
    """The tcgen05 / TMEM tile-GEMM lookup (rvo_corr_tiles) against the float64 oracle, through the
    tile-layout -> reference-layout index map; windows partly and entirely outside the maps included."""
    rng = np.random.default_rng(14)
    Np, Nf, C, H, W, P, E = 60, 5, 128, 30, 40, 3, 700
    gmap, pyr = synth.make_features(Nf, Np, C, H, W, P, seed=14)
    kk = rng.integers(0, 3 * Np, E)
    jj = rng.integers(0, 3 * Nf, E)
    _, _, coords = _corr_case(rng, E, Np, Nf, C, H, W, P, spread)
    coords[:5] = 1e9          # far outside
    coords[5:8] = -50.0
    coords[8] = np.nan
    g_t = torch.from_numpy(gmap).cuda().permute(0, 3, 1, 2)[None]
    p_t = [torch.from_numpy(p).cuda().permute(0, 3, 1, 2)[None] for p in pyr]
    c_t = torch.from_numpy(coords).cuda()[None]
    k_t, j_t = torch.from_numpy(kk).cuda(), torch.from_numpy(jj).cuda()
    out = altcorr.corr_tiles(g_t, p_t, c_t, k_t, j_t, Np, Nf)
    assert out.shape == (1, E, 1008) and out.dtype == torch.float16
    cols, ref = altcorr.tile_layout_index(2)
    got = torch.empty(E, 882, dtype=torch.float16, device="cuda")
    got[:, ref.cuda()] = out[0][:, cols.cuda()]
    exp = O.corr_pyramid(gmap.transpose(0, 3, 1, 2), [p.transpose(0, 3, 1, 2) for p in pyr],
                         np.nan_to_num(coords, nan=-1e9), kk, jj, Np, Nf, 3)
    g = got.float().cpu().numpy().astype(np.float64)
    assert np.isfinite(g).all()
    assert (np.abs(g - exp) <= 2.0 ** -10 * np.abs(exp) + 2e-4).all()
    # pad columns stay zero, and the per-edge kernel agrees
    pad = torch.ones(1008, dtype=torch.bool)
    pad[cols] = False
    assert (out[0][:, pad.cuda()] == 0).all()
    v1 = altcorr.corr_pyramid(g_t, p_t, torch.nan_to_num(c_t, nan=-1e9), k_t, j_t, Np, Nf, 3)
    assert (v1[0].float() - got.float()).abs().max().item() < 2e-3


def test_corr_tiles_full_size_agrees_with_per_edge_kernel():
    prob = synth.make_problem("default", 40, seed=1)
    gmap, pyr = synth.make_features(32, 96 * 32, seed=1)
    c = O.reproject(prob["poses"], prob["patches"], prob["intrinsics"], prob["ii"], prob["jj"],
                    prob["kk"]).astype(np.float32)
    g_t = torch.from_numpy(gmap).cuda().permute(0, 3, 1, 2)[None]
    p_t = [torch.from_numpy(p).cuda().permute(0, 3, 1, 2)[None] for p in pyr]
    c_t = torch.from_numpy(c).cuda()[None]
    k_t, j_t = torch.from_numpy(prob["kk"]).cuda(), torch.from_numpy(prob["jj"]).cuda()
    a = altcorr.corr_pyramid(g_t, p_t, c_t, k_t, j_t, 96 * 32, 32, 3)
    b = altcorr.corr_tiles(g_t, p_t, c_t, k_t, j_t, 96 * 32, 32)
    cols, ref = altcorr.tile_layout_index(2)
    assert (a[0][:, ref.cuda()].float() - b[0][:, cols.cuda()].float()).abs().max().item() < 2e-3


@pytest.mark.parametrize("case", ["one_tile_one_row", "rows_64", "rows_65", "rows_129", "long_chunks"])
def test_corr_tiles_block_structure_corners(case):
    """Inputs built to hit the corners of the tile kernel's block structure (corr_tc.cu): every row in ONE tile with
    ONE window-origin row (many 128-row blocks whose MMA covers only 8 of the 16 tile rows), tiles holding exactly
    64 / 65 rows (the limit of the blocks whose rows sit in two TMEM lanes and are split between two epilogue
    warps), 129 rows (a full block + a 1-row block of the same tile), and a graph large enough for the schedule
    with 8-block chunks (tile reuse across blocks).  Checked against the float64 oracle and the per-edge kernel."""
    rng = np.random.default_rng(31)
    Np, Nf, C, H, W, P = 60, 5, 128, 30, 40, 3
    E = {"one_tile_one_row": 100, "rows_64": 24, "rows_65": 24, "rows_129": 40, "long_chunks": 2400}[case]
    gmap, pyr = synth.make_features(Nf, Np, C, H, W, P, seed=31)
    kk = rng.integers(0, Np, E)
    jj = rng.integers(0, Nf, E)
    _, _, coords = _corr_case(rng, E, Np, Nf, C, H, W, P, 1.0)
    if case == "one_tile_one_row":
        jj[:] = 2
        coords[:, 0] = rng.uniform(14.0, 16.9, (E, P, P))       # window origins x0 = 11..13, one level-1 tile column
        coords[:, 1] = rng.uniform(17.01, 17.99, (E, P, P))     # y0 = 14 for every row: one window-origin row
    elif case in ("rows_64", "rows_65", "rows_129"):
        n_in = int(case.split("_")[1])
        jj[:] = 1
        coords[:] = 1e6                                          # everything else: outside the maps (zero rows)
        n = np.arange(n_in)
        e_, py, px = n // 9, (n % 9) // 3, n % 3
        coords[e_, 0, py, px] = rng.uniform(13.0, 19.9, n_in)    # tile x index 2 at level 1 (origins 10..16)
        coords[e_, 1, py, px] = rng.uniform(13.0, 19.9, n_in)    # all nine window-origin rows
    g_t = torch.from_numpy(gmap).cuda().permute(0, 3, 1, 2)[None]
    p_t = [torch.from_numpy(p).cuda().permute(0, 3, 1, 2)[None] for p in pyr]
    c_t = torch.from_numpy(coords).cuda()[None]
    k_t, j_t = torch.from_numpy(kk).cuda(), torch.from_numpy(jj).cuda()
    out = altcorr.corr_tiles(g_t, p_t, c_t, k_t, j_t, Np, Nf)
    cols, ref = altcorr.tile_layout_index(2)
    got = torch.empty(E, 882, dtype=torch.float16, device="cuda")
    got[:, ref.cuda()] = out[0][:, cols.cuda()]
    exp = O.corr_pyramid(gmap.transpose(0, 3, 1, 2), [p.transpose(0, 3, 1, 2) for p in pyr], coords, kk, jj, Np, Nf, 3)
    g = got.float().cpu().numpy().astype(np.float64)
    assert np.isfinite(g).all()
    assert (np.abs(g - exp) <= 2.0 ** -10 * np.abs(exp) + 2e-4).all()
    if case.startswith("rows_"):
        assert np.count_nonzero(np.abs(exp).reshape(E, -1).sum(-1)) > 0          # the rows inside the map are live
    v1 = altcorr.corr_pyramid(g_t, p_t, c_t, k_t, j_t, Np, Nf, 3)
    assert (v1[0].float() - got.float()).abs().max().item() < 2e-3
