"""CPU, world_size 2 over gloo: the host-side logic of the sharded patch graph.
The per-shard reduced camera systems (computed with the numpy oracle — the CUDA assemble kernel
needs a GPU) all-reduced by rampvo_b200.sharded.reduce_system equal the system of the whole graph,
and exchange_depths replicates the owners' depths."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_ops as O
from rampvo_b200 import sharded, synth
from tests.util import perturb_poses, targets_from_reprojection


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prob = synth.make_problem("cfg1", 8, seed=41)
        tgt = targets_from_reprojection(prob, O)
        prob["poses"] = perturb_poses(prob)
        prob["t0"] = 3
        mask = sharded.partition_edges(torch.from_numpy(prob["ii"]), world, rank).numpy()
        assert mask.any() and not mask.all()
        # ownership is by source frame: a patch never straddles ranks
        for k in np.unique(prob["kk"][mask]):
            assert mask[prob["kk"] == k].all()
        s = O.ba_system(prob["poses"], prob["patches"], prob["intrinsics"], tgt[mask], prob["weight"][mask],
                        1e-4, prob["ii"][mask], prob["jj"][mask], prob["kk"][mask], prob["t0"], prob["t1"])
        n6 = len(s["y"])
        Sy = torch.from_numpy(np.concatenate([s["S"], s["y"][:, None]], 1).copy())
        sharded.reduce_system(Sy)
        full = O.ba_system(prob["poses"], prob["patches"], prob["intrinsics"], tgt, prob["weight"], 1e-4,
                           prob["ii"], prob["jj"], prob["kk"], prob["t0"], prob["t1"])
        err_S = np.abs(Sy[:, :n6].numpy() - full["S"]).max() / np.abs(full["S"]).max()
        err_y = np.abs(Sy[:, n6].numpy() - full["y"]).max() / np.abs(full["y"]).max()
        # depths: every rank perturbs only its own frames, exchange makes the replicas equal
        M = prob["M"]
        patches = torch.from_numpy(prob["patches"].copy())
        frames = torch.arange(8).repeat_interleave(M)
        mine = sharded.owner_of_frame(frames, world) == rank
        patches[mine, 2] += 1.0 + rank
        sharded.exchange_depths(patches, 0, 8, M)
        exp = torch.from_numpy(prob["patches"].copy())
        for r in range(world):
            exp[sharded.owner_of_frame(frames, world) == r, 2] += 1.0 + r
        out[rank] = (float(err_S), float(err_y), bool(torch.allclose(patches, exp)))
    finally:
        dist.destroy_process_group()


def test_sharded_system_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        eS, ey, ok = out[r]
        assert eS < 1e-12 and ey < 1e-12 and ok


def test_partition_is_balanced_on_default_graph():
    prob = synth.make_problem("default", 40, seed=0)
    ii = torch.from_numpy(prob["ii"])
    for world in (2, 4, 8):
        counts = [int(sharded.partition_edges(ii, world, r).sum()) for r in range(world)]
        assert sum(counts) == prob["E"]
        assert max(counts) / (prob["E"] / world) < 1.35      # 22 source frames round-robin


def test_sharded_edge_rules_partition_the_reference_graph():
    """Ramp_vo(world_size=G): the forward / backward edges each rank appends (ramp/Ramp_vo.py:312-325 restricted to
    the frames it owns) are a partition of the single-GPU edge set, every edge sits on the owner of its source
    frame, and ownership follows a frame through a keyframe removal."""
    import types
    from rampvo_b200.Ramp_vo import Ramp_vo
    from rampvo_b200.config import preset
    cfg = preset("default")
    M = 4
    for world in (2, 3, 8):
        vos = [types.SimpleNamespace(cfg=cfg, M=M, n=0, counter=0, world_size=world, rank=r, _owner=[],
                                     device=torch.device("cpu")) for r in range(world)]
        full = types.SimpleNamespace(cfg=cfg, M=M, n=0, counter=0, world_size=1, rank=0, _owner=[],
                                     device=torch.device("cpu"))
        for t in range(20):
            for v in vos + [full]:
                v.counter += 1
                v._owner = v._owner[:v.n] + [(v.counter - 1) % v.world_size]
                v.n += 1
            want = set()
            for fn in (Ramp_vo._edges_forw, Ramp_vo._edges_back):
                kk, jj, pairs = fn(full)
                want |= set(zip(kk.tolist(), jj.tolist()))
                assert sum(pairs.values()) == kk.numel()
            got = []
            for v in vos:
                for fn in (Ramp_vo._edges_forw, Ramp_vo._edges_back):
                    kk, jj, pairs = fn(v)
                    assert sum(pairs.values()) == kk.numel()
                    assert all(v._owner[k // M] == v.rank for k in kk.tolist())
                    got += list(zip(kk.tolist(), jj.tolist()))
            assert len(got) == len(set(got)) and set(got) == want
            if t == 12:     # a keyframe drop renumbers the frames; the owner list shifts with them
                for v in vos + [full]:
                    del v._owner[v.n - 4]
                    v.n -= 1
