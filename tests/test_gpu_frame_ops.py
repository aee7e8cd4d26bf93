"""GPU parity of the per-frame glue kernels (csrc/frame_ops.cu) and of the fused BA target formation:
SURVEY.md rows a1 (Patchifier), a4 (patch selection), a8 (pyramid), a11 (filter_features), f2 (event stack)."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_gpu_vo as R
from oracle import ref_ops as O
from rampvo_b200 import _lib, fastba, synth, vo_utils
from tests import golden_inputs as GI
from tests.util import problem_tensors, targets_from_reprojection

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _torch_selection(events, M, border, nms):
    """the reference's op sequence (ramp/utils.py:186-226) with torch ops on the device"""
    ev = torch.abs(events.squeeze(0))
    ev = F.avg_pool2d(ev, 4, 4).transpose(3, 2)
    m = torch.mean(ev, dim=1)
    if border:
        m[:, :border, :] = 0
        m[:, -border:, :] = 0
        m[:, :, :border] = 0
        m[:, :, -border:] = 0
    if nms:
        m = vo_utils.nms_image(m, kernel_size=nms)
    flat = torch.flatten(m, start_dim=1)
    _, idx = torch.topk(flat, k=M, dim=-1)
    return torch.stack((idx / m.shape[-1], (idx % m.shape[-1]).float()), dim=-1)


# ------------------------------------------------------------------ a4: patch selection

def test_selection_matches_reference_fixture_bit_exact():
    z = np.load(os.path.join(G, "patch_selection.npz"))["coords"]
    ev = GI.selection_events("cuda")
    got = vo_utils.coords_from_topk_events(ev, 96, non_max_supp_rad=11)
    assert got.shape == (1, 96, 2) and got.dtype == torch.float32
    got = got.cpu().numpy()
    # the fixture was produced on the CPU, where idx / H' is a true division; on CUDA torch multiplies by the
    # reciprocal (<= 1 ulp apart).  Same cells, same order; y exact.
    assert (got[..., 1] == z[..., 1]).all()
    assert (np.rint(got[..., 0] * 120) == np.rint(z[..., 0] * 120)).all()
    assert np.abs(got[..., 0] - z[..., 0]).max() <= 2e-5


@pytest.mark.parametrize("M,border,nms,kind", [(96, 0, 11, "stream"), (32, 0, 11, "stream"), (300, 0, 11, "stream"),
                                               (96, 3, 11, "stream"), (96, 0, 0, "stream"), (96, 0, 5, "ties"),
                                               (48, 0, 11, "sparse"), (96, 0, 11, "constant")])
def test_selection_matches_torch_ops_and_oracle(M, border, nms, kind):
    """bit-exact against the reference's torch op sequence run on the same GPU (torch.topk's CUDA tie order) and
    against the numpy oracle; 'ties' = values from a 4-level alphabet, 'sparse' = fewer positives than M,
    'constant' = every cell equal (NMS keeps them all: pure index-order selection)"""
    g = torch.Generator(device="cuda").manual_seed(5)
    if kind == "stream":
        ev = synth.SyntheticSequence(seed=4, device="cuda").frame(3)[0]
    elif kind == "ties":
        ev = torch.randint(0, 2, (1, 1, 5, 480, 640), generator=g, device="cuda").float()
    elif kind == "sparse":
        ev = torch.zeros(1, 1, 5, 480, 640, device="cuda")
        pos = torch.randint(0, 480 * 640, (30,), generator=g, device="cuda")
        ev.view(5, -1)[2, pos] = torch.randint(1, 5, (30,), generator=g, device="cuda").float()
    else:
        ev = torch.full((1, 1, 5, 480, 640), 2.0, device="cuda")
    got = vo_utils.coords_from_topk_events(ev, M, border_suppression_size=border, non_max_supp_rad=nms)
    exp_t = _torch_selection(ev, M, border, nms)
    assert torch.equal(got, exp_t), "vs torch ops on CUDA"
    exp_o = O.select_patches(ev[0, 0].cpu().numpy(), M, border, nms, cuda_division=True)
    g = got[0].cpu().numpy()
    if M > 32:      # stable final sort: ties by ascending index, the oracle's rule
        assert (g == exp_o).all(), "vs oracle"
    else:           # k <= 32: torch's bitonic sort permutes ties; same cells as the oracle
        assert sorted(map(tuple, g.tolist())) == sorted(map(tuple, exp_o.tolist())), "vs oracle (as a set)"


def test_selection_matches_reference_function_on_gpu():
    """the reference's own get_coords_from_topk_events (staged ramp/utils.py) on the GPU"""
    if not R.available():
        pytest.skip("oracle/_ref not built")
    ns = R.load(with_vo=False)
    for seed in range(3):
        ev = synth.SyntheticSequence(seed=seed, device="cuda").frame(seed)[0]
        ref = ns.utils.get_coords_from_topk_events(events=ev, patches_per_image=96, border_suppression_size=0,
                                                   non_max_supp_rad=11)
        got = vo_utils.coords_from_topk_events(ev, 96, non_max_supp_rad=11)
        assert torch.equal(got, ref)


# ------------------------------------------------------------------ a8: pyramid level 2

@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_pyramid_level2_matches_avg_pool(dtype):
    g = torch.Generator(device="cuda").manual_seed(1)
    f = torch.randn(120, 160, 128, generator=g, device="cuda").to(dtype)
    got = vo_utils.pyramid_level2(f)
    ref = F.avg_pool2d(f.permute(2, 0, 1)[None], 4, 4)[0].permute(1, 2, 0)      # Ramp_vo.py:381
    assert got.shape == (30, 40, 128) and got.dtype == dtype
    exp = O.pyramid_level2(f.cpu().numpy())
    assert (got.cpu().numpy() == exp).all(), "vs oracle (fp32 accumulation, one rounding)"
    d = (got.float() - ref.float()).abs().max().item()
    assert d <= (1e-3 if dtype == torch.float16 else 1e-6), d   # torch may sum the window in another order


def test_copy_segments():
    a = [torch.randn(n, device="cuda") for n in (4, 1024, 300000, 96 * 27)]
    b = [torch.zeros_like(t) for t in a]
    vo_utils.copy_segments(list(zip(a, b)))
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    c = torch.arange(37, device="cuda", dtype=torch.uint8)      # odd size -> byte path
    d = torch.zeros_like(c)
    vo_utils.copy_segments([(c, d), (a[0], b[0])])
    assert torch.equal(c, d)


# ------------------------------------------------------------------ frame commit (Ramp_vo.py:345-372)

@pytest.mark.parametrize("initialized", [False, True])
def test_frame_commit_matches_the_reference_statements(initialized):
    g = torch.Generator(device="cuda").manual_seed(2)
    N, M, P, n = 16, 96, 3, 7
    patches_ = torch.rand(N, M, 3, P, P, generator=g, device="cuda")
    patches_[:, :, 2] = patches_[:, :, 2, :1, :1]               # one depth per patch, like the VO keeps it
    new = torch.rand(1, M, 3, P, P, generator=g, device="cuda")
    clr = torch.rand(1, M, 3, generator=g, device="cuda") * 2 - 0.5
    rnd = torch.rand(1, M, 1, 1, generator=g, device="cuda")
    intr = [320.0, 320.0, 320.0, 240.0]
    # reference statements
    exp_p = patches_.clone()
    pr = new.clone()
    pr[:, :, 2] = rnd
    if initialized:
        pr[:, :, 2] = torch.median(exp_p[n - 3:n, :, 2])
    exp_p[n] = pr
    exp_c = ((clr[0][:, [2, 1, 0]] + 0.5) * (255.0 / 2)).to(torch.uint8)
    # ours
    tst = torch.zeros(N, dtype=torch.long, device="cuda")
    K = torch.zeros(N, 4, device="cuda")
    index = torch.zeros(N, M, dtype=torch.long, device="cuda")
    imap = torch.zeros(N, dtype=torch.long, device="cuda")
    colors = torch.zeros(N, M, 3, dtype=torch.uint8, device="cuda")
    intr4 = (ctypes.c_float * 4)(*[v / 4 for v in intr])
    L = _lib.lib()
    _lib.check(L.rvo_frame_commit(_lib.ptr(new), _lib.ptr(clr), None if initialized else _lib.ptr(rnd.contiguous()),
                                  _lib.ptr(patches_), _lib.ptr(tst), _lib.ptr(K), _lib.ptr(index), _lib.ptr(imap),
                                  _lib.ptr(colors), intr4, n, M, P, N, 41, (n + 1) * M, 3 if initialized else 0,
                                  _lib.stream_ptr()), "rvo_frame_commit")
    assert torch.equal(patches_, exp_p)
    assert torch.equal(colors[n], exp_c)
    assert tst[n].item() == 41 and imap[n + 1].item() == (n + 1) * M
    assert (index[n + 1] == n + 1).all() and (index[n] == 0).all()
    assert torch.equal(K[n], torch.tensor(intr, device="cuda") / 4)


# ------------------------------------------------------------------ f2: event stack

def test_event_stack_matches_reference_fixture_bit_exact():
    from rampvo_b200.events import Events, EventToStack
    z = np.load(os.path.join(G, "event_stack.npz"))["stack"]
    x, y, p, ht, wd = GI.event_stream()
    ev = Events(x, y, np.zeros(len(x), np.int64), p, wd, ht)
    f32, i8 = EventToStack(5)(ev, return_int8=True)
    assert (i8.cpu().numpy() == z).all()
    assert (f32.cpu().numpy() == z.astype(np.float32)).all()


@pytest.mark.parametrize("n", [0, 1, 2, 777, 500000])
def test_event_stack_matches_oracle(n):
    from rampvo_b200.events import Events, EventToStack
    rng = np.random.default_rng(n)
    ht, wd = 480, 640
    x = rng.integers(0, wd, n).astype(np.uint16)
    y = rng.integers(0, ht, n).astype(np.uint16)
    p = (rng.integers(0, 2, n) * 2 - 1).astype(np.int8)
    got = EventToStack(5)(Events(x, y, np.zeros(n, np.int64), p, wd, ht))
    exp = O.event_stack(x, y, p, 5, ht, wd).astype(np.float32)
    assert got.shape == (5, ht, wd) and (got.cpu().numpy() == exp).all()


# ------------------------------------------------------------------ a11: filter_features + target inside BA

def test_ba_fused_target_equals_explicit_target_and_filter():
    """rvo_ba_forward_fused(coords, delta, weight) == filter_features + fastba.BA(target, weight) bit for bit,
    and the filtered confidences equal the reference's filter_features (ramp/utils.py:557-570)."""
    prob = synth.make_problem("default", 40, seed=3)
    t = problem_tensors(prob)
    E = prob["E"]
    g = torch.Generator(device="cuda").manual_seed(9)
    from rampvo_b200 import projective_ops as pops
    from rampvo_b200.lietorch import SE3
    coords = pops.reproject_cf(SE3(t["poses"]), t["patches"], t["intrinsics"], t["ii"], t["jj"], t["kk"])
    delta = torch.randn(1, E, 2, generator=g, device="cuda") * 30          # pushes many targets out of the image
    weight = torch.rand(1, E, 2, generator=g, device="cuda")
    lm = torch.tensor([1e-4], device="cuda")
    # explicit path (Ramp_vo.py:288-304)
    target = coords[..., 1, 1] + delta
    wf = vo_utils.filter_features(confidences=weight, target=target, data_shape=(120, 160))
    assert 0.02 < (wf == 0).float().mean().item() < 0.98
    p1, q1 = t["poses"].clone(), t["patches"].clone()
    fastba.BA(p1, q1, t["intrinsics"], target, wf, lm, t["ii"], t["jj"], t["kk"], prob["t0"], prob["t1"],
              prob["M"], 2)
    # fused path
    from rampvo_b200.net import GraphPlans
    plans = GraphPlans(t["ii"], t["jj"], t["kk"])
    p2, q2 = t["poses"].clone(), t["patches"].clone()
    wout = torch.empty_like(weight)
    t0d = torch.tensor([prob["t0"]], dtype=torch.int32, device="cuda")
    fastba.BA_fused(p2, q2, t["intrinsics"], coords, delta, weight, 120, 160, lm, t["ii"], t["jj"], plans.plan_k,
                    prob["t1"] - prob["t0"], t0d, 2, weight_out=wout)
    assert torch.equal(wout, wf)
    if R.available():
        ns = R.load(with_vo=False)
        assert torch.equal(wout, ns.utils.filter_features(confidences=weight, target=target, data_shape=(120, 160)))
    # the pose blocks are accumulated with shared-memory atomics: agreement to accumulation-order noise
    assert (p1 - p2).abs().max().item() < 1e-5 * p1.abs().max().item()
    assert (q1 - q2).abs().max().item() < 1e-5 * q1.abs().max().item()


# ------------------------------------------------------------------ a1: whole Patchifier vs the reference's

def test_patchifier_matches_reference_patchifier_on_gpu():
    """network.patchify (encoder -> /4 -> selection -> 4 patch gathers) against the reference's own Patchifier
    (ramp/net.py:128-203) on the same GPU with the same weights, 480x640, under autocast like Ramp_vo.py:331."""
    if not R.available():
        pytest.skip("oracle/_ref not built")
    from rampvo_b200.net import VONet
    cfg = {"event_bias": True, "input_mode": "MultiScale", "num_event_bins": 5}
    torch.manual_seed(1234)
    mine = VONet(cfg).cuda().eval()
    ns = R.load(with_vo=False)
    ref = ns.net.VONet(cfg)
    ref.load_state_dict(mine.state_dict(), strict=True)
    ref = ref.cuda().eval()
    seq = synth.SyntheticSequence(seed=2, device="cuda")
    with torch.no_grad(), torch.autocast("cuda", enabled=True):
        for t in range(2):                      # the second frame exercises the carried super state
            ev, im, mask = seq.frame(t)
            fr, gr, ir, pr, xr, cr = ref.patchify(input_=(ev, im, mask), patches_per_image=96, event_bias=True,
                                                  reinit_hidden=(t == 0))
            fo, go, io, po, xo, co = mine.patchify(input_=(ev, im, mask), patches_per_image=96, event_bias=True,
                                                   reinit_hidden=(t == 0))
            assert torch.equal(po, pr), "patch coordinates / grid gather are exact"
            assert torch.equal(xo, xr)
            assert (co.float() - cr.float()).abs().max().item() < 1e-6
            scale = fr.float().abs().max().item()
            for name, a, b in (("fmap", fo, fr), ("gmap", go, gr), ("imap", io, ir)):
                d = (a.float() - b.float().view_as(a)).abs().max().item()
                print("[patchifier frame %d] %s max|d|/max|fmap| = %.3e" % (t, name, d / scale))
                assert d < 3e-2 * scale, name       # fp16 convolutions on both sides, different summation orders


# ------------------------------------------------------------------ a3: SingleScale encoder (BASELINE configs[0])

def test_single_scale_encoder_fused_path_matches_reference_fixture():
    """rvo_scene_lstm_forward (carried per-pixel LSTM state, device-side presence flags) + channels-last CNNs vs
    the fixture generated by the reference's MergerLSTMsceneEncoder (fp32 CPU)."""
    from rampvo_b200.extractor import MergerLSTMsceneEncoder
    z = np.load(os.path.join(G, "single_scale_encoder.npz"))
    torch.manual_seed(GI.ENCODER_SEED)
    enc = MergerLSTMsceneEncoder(5, 3).cuda().eval()
    with torch.no_grad(), torch.autocast("cuda", enabled=True):
        for f, (ev, im) in enumerate(GI.single_scale_inputs("cuda")):
            fmap, imap, _ = enc(events=ev, images=im, reinit_hidden=(f == 0))
            assert fmap.dtype == torch.float16 and fmap.shape == (1, 1, 128, 8, 12)
            for name, got in (("fmap", fmap), ("imap", imap)):
                ref = z["%s_%d" % (name, f)]
                d = np.abs(got[0, 0].float().cpu().numpy() - ref).max() / np.abs(ref).max()
                print("[single-scale frame %d] %s rel err %.3e" % (f, name, d))
                assert d < 3e-2, (name, f, d)       # fp16 convolutions vs the fp32 reference


def test_single_scale_vo_tracks_the_reference_vo():
    """BASELINE.json configs[0] on the GPU: SingleScale encoder, 32 patches, 8+ frames — ours vs the reference's
    own Ramp_vo with the same weights: identical patch graph, poses compared after the initialisation."""
    if not R.available():
        pytest.skip("oracle/_ref not built")
    from rampvo_b200.Ramp_vo import Ramp_vo
    from rampvo_b200.config import preset
    from rampvo_b200.net import VONet
    cfg_t = {"event_bias": True, "input_mode": "SingleScale", "num_event_bins": 5}
    torch.manual_seed(1234)
    net = VONet(cfg_t)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    cfg = preset("cfg1")
    cfg.KEYFRAME_THRESH = 0.0
    ref = R.make_vo(cfg.clone(), sd, cfg_t)
    ours = Ramp_vo(cfg.clone(), net, cfg_t, device="cuda")
    ref.motion_probe = lambda: torch.tensor(10.0)
    ours.motion_probe = lambda: torch.tensor(10.0)
    seq = synth.SyntheticSequence(seed=3, device="cuda")
    intr = seq.intrinsics.cuda()
    with torch.no_grad():
        for t in range(10):
            fr = seq.frame(t)
            torch.manual_seed(100 + t)
            ref(t, fr, intr)
            torch.manual_seed(100 + t)
            ours(t, fr, intr)
            ours.sync()
            assert ref.n == ours.n and torch.equal(ref.ii, ours.ii) and torch.equal(ref.kk, ours.kk)
            assert torch.equal(ref.patches_[:ref.n, :, :2], ours.patches_[:ours.n, :, :2]), "patch selection"
    assert ref.is_initialized and ours.is_initialized
    a, b = ref.poses_[:ref.n], ours.poses_[:ours.n]
    ext = max(float((a[:, :3] - a[:1, :3]).norm(dim=-1).max()), 1e-3)
    dt = float((a[:, :3] - b[:, :3]).norm(dim=-1).max()) / ext
    print("[single-scale VO] pose dt/extent after %d frames: %.3e" % (ref.n, dt))
    assert np.isfinite(dt) and dt < 0.2
