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
