"""END-TO-END differential tests against THE REFERENCE ITSELF running on the same GPU.

oracle/ref_gpu_vo.py imports the reference's own, unmodified `ramp` package (Ramp_vo state machine, VONet,
Update, extractor, projective_ops, lietorch wrappers) on top of the reference's CUDA ops compiled from
/root/reference (cuda_corr, cuda_ba).  These tests drive that object and rampvo_b200.Ramp_vo with the same
weights, the same frames and the reference's own call order (ramp/Ramp_vo.py:327-410 -> :276-310) and compare

  * the patch graph (ii / jj / kk, n, m) — exact,
  * one recurrent update from an IDENTICAL state (teacher forcing): hidden state, confidence weights, poses
    and depths after reproject -> corr -> Update -> 2 BA iterations,
  * the free-running trajectories (both implementations accumulate their own fp16 rounding).

Tolerances are stated where they are asserted; measured values are printed (-s) and recorded in
profiles/r02_e2e_parity.txt.
"""
import numpy as np
import pytest
import torch

from oracle import ref_gpu_vo as R
from rampvo_b200 import synth
from rampvo_b200.config import preset

pytestmark = pytest.mark.gpu

TRAIN_CFG = {"event_bias": True, "input_mode": "MultiScale", "num_event_bins": 5}


def _need():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    if not R.available():
        pytest.skip("oracle/_ref (compiled reference ops + staged reference python) is not built")


def _pair(cfg_name, thresh, use_graphs, seed=1234, ht=480, wd=640):
    """(reference Ramp_vo, ours) with identical weights and the data-dependent init gate pinned on both
    (random weights never move the probe; bench.py pins it the same way)."""
    from rampvo_b200.Ramp_vo import Ramp_vo
    from rampvo_b200.net import VONet
    torch.manual_seed(seed)
    net = VONet(TRAIN_CFG)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    cfg = preset(cfg_name)
    cfg.KEYFRAME_THRESH = thresh
    ref = R.make_vo(cfg.clone(), sd, TRAIN_CFG, ht=ht, wd=wd)
    ours = Ramp_vo(cfg.clone(), net, TRAIN_CFG, ht=ht, wd=wd, device="cuda", use_graphs=use_graphs)
    ref.motion_probe = lambda: torch.tensor(10.0)
    ours.motion_probe = lambda: torch.tensor(10.0)
    return ref, ours


def _frames(n, seed=3):
    seq = synth.SyntheticSequence(seed=seed, device="cuda")
    return [seq.frame(t) for t in range(n)], seq.intrinsics.cuda()


def _graph_equal(ref, ours):
    ours.sync()
    return (ref.n == ours.n and ref.m == ours.m and torch.equal(ref.ii, ours.ii) and torch.equal(ref.jj, ours.jj)
            and torch.equal(ref.kk, ours.kk))


def _pose_err(ref, ours):
    """max translation difference relative to the trajectory extent, max rotation difference (rad)"""
    n = ref.n
    a, b = ref.poses_[:n].double(), ours.poses_[:n].double()
    ext = max(float((a[:, :3] - a[:1, :3]).norm(dim=-1).max()), 1e-3)
    dt = float((a[:, :3] - b[:, :3]).norm(dim=-1).max()) / ext
    # rotation angle of q_a * conj(q_b) from the norm of its vector part (well conditioned near identity,
    # unlike acos of the dot product of two fp32-normalised quaternions)
    qa = a[:, 3:] / a[:, 3:].norm(dim=-1, keepdim=True)
    qb = b[:, 3:] / b[:, 3:].norm(dim=-1, keepdim=True)
    va, wa, vb, wb = qa[:, :3], qa[:, 3:], qb[:, :3], qb[:, 3:]
    vec = wb * va - wa * vb - torch.cross(va, vb, dim=-1)
    return dt, float((2 * torch.asin(vec.norm(dim=-1).clamp(max=1.0))).max())


def _copy_state(ref, ours):
    """make `ours` hold exactly the reference's state (buffers are re-laid-out, not re-computed)"""
    ours.sync()
    ours.n, ours.m, ours.counter, ours.is_initialized = ref.n, ref.m, ref.counter, ref.is_initialized
    ours.tlist = list(ref.tlist)
    for name in ("tstamps_", "poses_", "patches_", "intrinsics_", "index_", "index_map_", "colors_", "imap_"):
        getattr(ours, name).copy_(getattr(ref, name))
    mem, M, P = ref.mem, ref.M, ref.P
    ours._gmap_store.copy_(ref.gmap_.view(mem * M, 128, P, P).permute(0, 2, 3, 1))
    ours._fmap1_store.copy_(ref.fmap1_[0].permute(0, 2, 3, 1))
    ours._fmap2_store.copy_(ref.fmap2_[0].permute(0, 2, 3, 1))
    ours.ii, ours.jj, ours.kk = ref.ii.clone(), ref.jj.clone(), ref.kk.clone()
    ours._pair_cnt, ours._plans = None, None
    E = ref.ii.numel()
    ours.net = torch.zeros(1, 0, ours.DIM, device=ours.device)
    ours._net_bufs, ours._net_cur = [None, None], 0
    ours._net_reserve(E)
    ours._net_bufs[0][:, :E] = ref.net.float()
    ours.net = ours._net_bufs[0][:, :E]
    ours._ugraphs.clear()


@pytest.mark.parametrize("cfg_name,n_frames", [("cfg1", 12), ("default", 14)])
def test_one_update_from_the_reference_state_matches(cfg_name, n_frames):
    """Teacher forcing: the reference tracks n_frames; its whole state is copied into ours; BOTH run one
    update() (reproject -> corr x2 levels -> Update -> filter -> 2 BA iterations -> point cloud)."""
    _need()
    ref, ours = _pair(cfg_name, 0.0, use_graphs=False)
    frames, intr = _frames(n_frames)
    with torch.no_grad():
        torch.manual_seed(7)
        for t in range(n_frames):
            ref(t, frames[t], intr)
        assert ref.is_initialized and ref.ii.numel() > 0
        _copy_state(ref, ours)
        p0 = ref.poses_[:ref.n].clone()
        ref.update()
        ours.update()
    E = ref.ii.numel()
    net_r, net_o = ref.net[0].float(), ours.net[0].float()
    scale = float(net_r.abs().max())
    d_net = float((net_r - net_o).abs().max()) / scale
    d_w = float((ref.last_weight - ours.last_weight).abs().max())
    moved = float((ref.poses_[:ref.n] - p0).abs().max())
    dt, dq = _pose_err(ref, ours)
    dd = float((ref.patches_[:ref.n, :, 2] - ours.patches_[:ref.n, :, 2]).abs().max())
    dpts = float((ref.points_[:ref.m] - ours.points_[:ref.m]).abs().max() /
                 ref.points_[:ref.m].abs().max().clamp(min=1e-6))
    print("\n[e2e one update, %s] E=%d  net max|d|/max|net| %.3e  weight max|d| %.3e  BA moved poses by %.3e  "
          "pose dt/extent %.3e  dq %.3e rad  depth max|d| %.3e  points rel %.3e"
          % (cfg_name, E, d_net, d_w, moved, dt, dq, dd, dpts))
    # hidden state / weights: both sides run the 21 Linear layers in fp16 (autocast) — the reference accumulates
    # the correlation dot products in fp16 as well (correlation_kernel.cu:121-130), we accumulate in fp32.
    # (measured, profiles/r02_e2e_parity.txt: net 6e-4 / 8e-4, weight 4.9e-4 = one fp16 ulp of a sigmoid output)
    assert d_net < 5e-3
    assert d_w < 2e-3
    # poses after 2 Gauss-Newton iterations driven by those fp16 weights / targets (measured: dt 1.6e-5 / 2.8e-5
    # of the trajectory extent, dq 1e-6 / 2e-6 rad, depths 6e-4 / 1.2e-3)
    assert dt < 5e-4 and dq < 5e-5
    assert dd < 1e-2


@pytest.mark.parametrize("cfg_name", ["cfg1", "default"])
def test_free_run_tracks_the_reference(cfg_name):
    """Both state machines consume the same 16 frames independently (ours through its CUDA graphs, the product
    path).  The patch graph must be identical at every frame; poses are compared right after the initialisation
    (frame 8: 12 updates) and the drift afterwards is reported."""
    _need()
    ref, ours = _pair(cfg_name, 0.0, use_graphs=True)
    frames, intr = _frames(16)
    errs = []
    with torch.no_grad():
        for t in range(16):
            torch.manual_seed(100 + t)          # depth init draws torch.rand_like (Ramp_vo.py:367)
            ref(t, frames[t], intr)
            torch.manual_seed(100 + t)
            ours(t, frames[t], intr)
            assert _graph_equal(ref, ours), "patch graph diverged at frame %d" % t
            if ref.is_initialized:
                errs.append((t,) + _pose_err(ref, ours))
    assert errs, "the reference never initialised"
    print("\n[e2e free run, %s] frame: pose dt/extent, dq(rad): %s"
          % (cfg_name, "  ".join("%d: %.2e %.2e" % e for e in errs)))
    t, dt, dq = errs[0]
    assert np.isfinite([e[1] for e in errs]).all()
    if cfg_name == "default":
        # measured (profiles/r02_e2e_parity.txt): dt 5e-3 right after the initialisation, 2e-3 afterwards; dq 4e-5
        assert dt < 3e-2 and dq < 1e-3, "after the 12 initialisation updates (frame %d)" % t
        assert max(e[1] for e in errs) < 3e-2 and max(e[2] for e in errs) < 1e-3
    else:
        # cfg1 (32 patches, every pose but the first free, random weights): the 12 chained initialisation updates
        # amplify the fp16 differences along the gauge directions of a barely constrained window (the one-update
        # test above bounds a single step at 2e-5).  Measured: dt 1.4e-2, dq 1e-3.
        assert max(e[1] for e in errs) < 0.15 and max(e[2] for e in errs) < 1e-2
    assert torch.equal(ref.tstamps_[:ref.n], ours.tstamps_[:ours.n])


def test_keyframe_drops_follow_the_reference():
    """default.yaml with its real KEYFRAME_THRESH = 15: slow synthetic motion makes the reference drop keyframes
    (ramp/Ramp_vo.py:237-274); ours must drop the same frames, renumber the same edges and keep the same
    trajectory bookkeeping (tstamps_, delta keys) for as long as the decisions agree."""
    _need()
    ref, ours = _pair("cfg1", 15.0, use_graphs=True)
    frames, intr = _frames(20)
    agree = 0
    with torch.no_grad():
        for t in range(20):
            torch.manual_seed(100 + t)
            ref(t, frames[t], intr)
            torch.manual_seed(100 + t)
            ours(t, frames[t], intr)
            if not _graph_equal(ref, ours):
                break
            agree = t + 1
    drops = ref.counter - ref.n
    print("\n[e2e keyframes] frames agreeing %d/20, reference dropped %d keyframes, n=%d" % (agree, drops, ref.n))
    assert agree >= 12, "graphs diverged at frame %d" % agree
    if agree == 20:
        assert sorted(ref.delta.keys()) == sorted(ours.delta.keys())
        assert torch.equal(ref.tstamps_[:ref.n], ours.tstamps_[:ours.n])
        pr, _ = ref.terminate()
        po, _ = ours.terminate()
        assert pr.shape == po.shape
