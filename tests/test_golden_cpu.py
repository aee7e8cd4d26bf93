"""CPU: the oracle and the host-side modules against fixtures produced by the reference's own Python
code (tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import ref_ops as O
from tests import golden_inputs as GI
from tests.util import rel_err

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_transform_matches_reference_pops():
    z = np.load(os.path.join(G, "pops_transform.npz"))
    prob = GI.pops_problem()
    a = (prob["poses"], prob["patches"], prob["intrinsics"], prob["ii"], prob["jj"], prob["kk"])
    c, v, (Ji, Jj, Jz) = O.transform(*a, jacobian=True)
    np.testing.assert_allclose(c, z["coords"], rtol=0, atol=1e-9)
    assert (v == z["valid"]).all()
    for got, key in ((Ji, "Ji"), (Jj, "Jj"), (Jz, "Jz")):
        np.testing.assert_allclose(got, z[key], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(O.transform(*a, tonly=True)[0], z["coords_tonly"], atol=1e-9)
    np.testing.assert_allclose(O.flow_mag(*a, beta=0.5), z["flow_mag"], atol=1e-9)
    ix = np.arange(prob["patches"].shape[0]) // prob["M"]
    np.testing.assert_allclose(O.point_cloud_centers(prob["poses"], prob["patches"], prob["intrinsics"], ix),
                               z["points"], rtol=1e-9, atol=1e-9)


def test_oracle_ba_matches_reference_python_ba():
    """ramp/ba.py (ep=1, first 4 poses fixed) and the cuda_ba restatement are the same solver on an
    in-bounds graph (SURVEY.md section 8c lists where they differ: gates, clamps, damping)."""
    z = np.load(os.path.join(G, "python_ba.npz"))
    prob, tgt = GI.ba_problem(O)
    for it in (1, 2):
        p, q = O.ba(prob["poses"], prob["patches"], prob["intrinsics"], tgt, prob["weight"], 1e-4,
                    prob["ii"], prob["jj"], prob["kk"], prob["t0"], prob["t1"], iterations=it)
        # ramp/ba.py goes through lietorch (normalised quaternions), cuda_ba does not: 1e-6
        assert rel_err(p, z["poses_%d" % it]) < 1e-6
        assert rel_err(q[:, 2, 0, 0], z["disps_%d" % it]) < 1e-6


def test_encoder_matches_reference_on_cpu():
    """rampvo_b200.extractor (restructured forward, same parameter tree) vs ramp/extractor.py, fp32."""
    from rampvo_b200.extractor import MultiScaleMergerDoubleNet
    z = np.load(os.path.join(G, "encoder.npz"))
    torch.manual_seed(GI.ENCODER_SEED)
    enc = MultiScaleMergerDoubleNet(5, 3).eval()
    frames = GI.encoder_inputs()
    with torch.no_grad():
        for f, (ev, im) in enumerate(frames):
            fmap, imap = enc(events=ev, images=im, mask=torch.tensor([True]), reinit_hidden=(f == 0))
            assert fmap.shape == (1, 1, 128, 16, 24) and imap.shape == (1, 1, 384, 16, 24)
            assert rel_err(fmap[0, 0].numpy(), z["fmap_%d" % f]) < 1e-4
            assert rel_err(imap[0, 0].numpy(), z["imap_%d" % f]) < 1e-4
        enc(events=frames[0][0], images=frames[0][1], mask=torch.tensor([False]))
        fmap, _ = enc(events=frames[1][0], images=frames[1][1], mask=torch.tensor([True]))
        assert rel_err(fmap[0, 0].numpy(), z["fmap_after_events_only"]) < 1e-4


def test_patch_selection_matches_reference():
    from rampvo_b200.vo_utils import coords_from_topk_events
    z = np.load(os.path.join(G, "patch_selection.npz"))
    c = coords_from_topk_events(GI.selection_events(), 96, non_max_supp_rad=11)
    assert c.shape == (1, 96, 2)
    assert (c.numpy() == z["coords"]).all()
    assert (c[..., 0] != c[..., 0].floor()).any()       # the fractional-x quirk of utils.py:212


def test_state_dict_keys_match_reference_checkpoint_layout():
    from rampvo_b200.net import VONet
    net = VONet({"event_bias": True, "input_mode": "MultiScale", "num_event_bins": 5})
    keys = set(net.state_dict().keys())
    for k in ("patchify.encoder.ev_encoders.2.convlstm.weight_hh_l0",
              "patchify.encoder.super_state_im_encoders.0.encoder.bias",
              "patchify.encoder.fmap_encoder.layer2.0.downsample.0.weight",   # dead weights kept
              "patchify.encoder.imap_encoder.conv2.bias",
              "patchify.encoder.fmap_encoder.layer3.0.downsample.0.weight",
              "update.corr.0.weight", "update.corr.3.bias", "update.gru.1.res.2.weight",
              "update.agg_ij.h.bias", "update.c2.2.weight", "update.d.1.weight", "update.w.1.bias"):
        assert k in keys, k
    n_enc = sum(v.numel() for k, v in net.state_dict().items() if k.startswith("patchify"))
    assert n_enc == 844766


def test_oracle_network_restatement_matches_reference_fixtures():
    """oracle/ref_vo.py (the CPU baseline path: nn.LSTM plumbing, index-op scatter softmax) against
    the reference's Update / encoder outputs."""
    from oracle import ref_vo
    from rampvo_b200.extractor import MultiScaleMergerDoubleNet
    from rampvo_b200.net import Update
    z = np.load(os.path.join(G, "update_op.npz"))
    torch.manual_seed(GI.UPDATE_SEED)
    p = {k: v.detach() for k, v in Update(3).state_dict().items()}
    g = GI.update_inputs()
    net, d, w = ref_vo.update_operator(p, g["net"][0], g["inp"][0], g["corr"][0], g["ii"], g["jj"], g["kk"])
    assert rel_err(net.numpy(), z["net"]) < 1e-5
    assert rel_err(d.numpy(), z["delta"]) < 1e-4 and rel_err(w.numpy(), z["weight"]) < 1e-5

    z = np.load(os.path.join(G, "encoder.npz"))
    torch.manual_seed(GI.ENCODER_SEED)
    sd = {"patchify.encoder." + k: v for k, v in MultiScaleMergerDoubleNet(5, 3).state_dict().items()}
    enc = ref_vo.Encoder(sd)
    for f, (ev, im) in enumerate(GI.encoder_inputs()):
        fmap, imap = enc(ev, im, [True], reinit_hidden=(f == 0))
        assert rel_err(fmap[0, 0].numpy(), z["fmap_%d" % f]) < 1e-5
        assert rel_err(imap[0, 0].numpy(), z["imap_%d" % f]) < 1e-5


def test_oracle_torch_corr_matches_numpy_corr():
    from oracle import ref_vo
    from rampvo_b200 import synth
    rng = np.random.default_rng(3)
    gmap, pyr = synth.make_features(4, 24, C=32, ht=24, wd=32, seed=3, dtype=np.float32)
    E = 60
    kk, jj = rng.integers(0, 24, E), rng.integers(0, 4, E)
    coords = rng.uniform(-4, 36, (E, 2, 1, 1)).astype(np.float32) + rng.normal(0, 0.6, (E, 2, 3, 3)).astype(np.float32)
    g = gmap.transpose(0, 3, 1, 2)
    p = [x.transpose(0, 3, 1, 2) for x in pyr]
    exp = O.corr_pyramid(g, p, coords, kk, jj, 0, 0, 3)
    got = ref_vo.corr_pyramid_torch(torch.from_numpy(g.copy()), [torch.from_numpy(x.copy()) for x in p],
                                    torch.from_numpy(coords), torch.from_numpy(kk), torch.from_numpy(jj),
                                    chunk=17)
    assert np.abs(got.numpy() - exp).max() < 1e-5


def test_oracle_event_stack_matches_reference_fixture():
    """oracle.event_stack vs the reference's EventToStack_Numpy output (int8, incl. wrapped hot cells)"""
    z = np.load(os.path.join(G, "event_stack.npz"))["stack"]
    x, y, p, ht, wd = GI.event_stream()
    got = O.event_stack(x, y, p, 5, ht, wd)
    assert got.dtype == z.dtype == np.int8 and (got == z).all()
    assert z.min() < -30       # the hot cell collects ~480 (+1) events per bin: the int8 cast wrapped it negative


def test_oracle_selection_matches_reference_fixture():
    """oracle.select_patches (value descending, ties by ascending index) vs get_coords_from_topk_events"""
    z = np.load(os.path.join(G, "patch_selection.npz"))["coords"]
    ev = GI.selection_events()
    got = O.select_patches(ev[0, 0].numpy(), 96, 0, 11)
    assert got.dtype == np.float32 and (got == z[0]).all()


def test_single_scale_encoder_matches_reference_on_cpu():
    """MergerLSTMsceneEncoder (SingleScale, BASELINE.json configs[0]) vs ramp/extractor.py:187-269: carried per-pixel
    LSTM state over three calls, absent image on the second"""
    from rampvo_b200.extractor import MergerLSTMsceneEncoder
    z = np.load(os.path.join(G, "single_scale_encoder.npz"))
    torch.manual_seed(GI.ENCODER_SEED)
    enc = MergerLSTMsceneEncoder(5, 3).eval()
    with torch.no_grad():
        for f, (ev, im) in enumerate(GI.single_scale_inputs()):
            fmap, imap, _ = enc(events=ev, images=im, reinit_hidden=(f == 0))
            assert rel_err(fmap[0, 0].numpy(), z["fmap_%d" % f]) < 1e-5
            assert rel_err(imap[0, 0].numpy(), z["imap_%d" % f]) < 1e-5


def test_differentiable_ba_matches_reference_python_ba():
    """rampvo_b200.ba.BA (the training-time Gauss-Newton step, tensor algebra) vs the fixture produced by the
    reference's ramp/ba.py — and gradients reach the confidence weights and targets through both iterations"""
    from rampvo_b200 import ba as myba
    from rampvo_b200.lietorch import SE3
    z = np.load(os.path.join(G, "python_ba.npz"))
    prob, tgt = GI.ba_problem(O)
    t = GI.as_torch(prob, torch.float64)
    poses, patches = SE3(t["poses"].clone().requires_grad_(True)), t["patches"].clone().requires_grad_(True)
    tg = torch.from_numpy(tgt).double()[None].requires_grad_(True)
    wg = torch.from_numpy(prob["weight"]).double()[None].requires_grad_(True)
    for it in (1, 2):
        poses, patches = myba.BA(poses, patches, t["intrinsics"], tg, wg, 1e-4, t["ii"], t["jj"], t["kk"],
                                 bounds=[-64, -64, 160 + 64, 120 + 64], ep=1.0, fixedp=prob["t0"])
        assert np.abs(poses.data[0].detach().numpy() - z["poses_%d" % it]).max() < 1e-12
        assert np.abs(patches[0, :, 2, 0, 0].detach().numpy() - z["disps_%d" % it]).max() < 1e-12
    (poses.data[..., :3].sum() + patches[:, :, 2].sum()).backward()
    assert torch.isfinite(wg.grad).all() and wg.grad.abs().sum() > 0
    assert torch.isfinite(tg.grad).all() and tg.grad.abs().sum() > 0


def test_encoder_clip_matches_reference_on_cpu():
    """T = 4 frames in ONE call: the reference's per-pixel nn.LSTM runs over the 4-step sequence
    (ramp/extractor.py:364-381), which only the training unroll exercises"""
    from rampvo_b200.extractor import MultiScaleMergerDoubleNet
    z = np.load(os.path.join(G, "encoder_clip.npz"))
    torch.manual_seed(GI.ENCODER_SEED)
    enc = MultiScaleMergerDoubleNet(5, 3).eval()
    ev, im, mask = GI.encoder_clip_inputs()
    with torch.no_grad():
        fmap, imap = enc(events=ev, images=im, mask=mask, reinit_hidden=True)
    assert fmap.shape == (1, 4, 128, 8, 12)
    assert rel_err(fmap[0].numpy(), z["fmap"]) < 1e-4
    assert rel_err(imap[0, :, :16].numpy(), z["imap16"]) < 1e-4
