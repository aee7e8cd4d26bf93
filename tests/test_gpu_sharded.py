"""GPU, world_size 2 over NCCL (skipped on a single-GPU box): fastba over a graph sharded by source
frame with the all-reduce of [S | y] equals the single-GPU solve."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from oracle import ref_ops as O
    from rampvo_b200 import fastba, sharded, synth
    from tests.util import perturb_poses, problem_tensors, targets_from_reprojection
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        prob = synth.make_problem("default", 40, seed=51)
        tgt = targets_from_reprojection(prob, O)
        prob["poses"] = perturb_poses(prob)
        dev = "cuda:%d" % rank
        t = problem_tensors(prob, dev)
        tg = torch.from_numpy(tgt).to(dev)
        wg = torch.from_numpy(prob["weight"]).to(dev)
        lm = torch.tensor([1e-4], device=dev)
        mask = sharded.partition_edges(t["ii"], world, rank)
        sharded.sharded_BA(t["poses"], t["patches"], t["intrinsics"], tg[mask][None], wg[mask][None], lm,
                           t["ii"][mask], t["jj"][mask], t["kk"][mask], prob["t0"], prob["t1"], 2)
        sharded.exchange_depths(t["patches"][0], 0, prob["n"], prob["M"])
        ref = problem_tensors(prob, dev)
        fastba.BA(ref["poses"], ref["patches"], ref["intrinsics"], tg[None], wg[None], lm, ref["ii"], ref["jj"],
                  ref["kk"], prob["t0"], prob["t1"], prob["M"], 2)
        ep = (t["poses"] - ref["poses"]).abs().max().item() / ref["poses"].abs().max().item()
        ed = (t["patches"][0, :, 2] - ref["patches"][0, :, 2]).abs().max().item() / ref["patches"][0, :, 2].abs().max().item()
        gathered = [torch.zeros_like(t["poses"]) for _ in range(world)]
        dist.all_gather(gathered, t["poses"])
        same = all(bool((g == gathered[0]).all()) for g in gathered)
        out[rank] = (ep, ed, same)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
def test_sharded_ba_matches_single_gpu():
    import torch.multiprocessing as mp
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        ep, ed, same = out[r]
        assert ep < 1e-5 and ed < 1e-5, (ep, ed)
        assert same            # every rank holds bit-identical poses after the replicated solve


def _vo_worker(rank, world, port, out):
    import torch.distributed as dist
    from rampvo_b200 import synth
    from rampvo_b200.Ramp_vo import Ramp_vo
    from rampvo_b200.config import preset
    from rampvo_b200.net import VONet
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        tc = {"event_bias": True, "input_mode": "MultiScale", "num_event_bins": 5}
        cfg = preset("default")
        cfg.KEYFRAME_THRESH = 0.0

        def make(ws, rk):
            torch.manual_seed(1234)
            vo = Ramp_vo(cfg.clone(), VONet(tc), tc, device=dev, world_size=ws, rank=rk)
            vo.motion_probe = lambda: torch.tensor(10.0)
            return vo
        sh, single = make(world, rank), make(1, 0)
        seq = synth.SyntheticSequence(seed=3, device=dev)
        with torch.no_grad():
            for t in range(14):
                fr = seq.frame(t)
                torch.manual_seed(100 + t)
                sh(t, fr, seq.intrinsics)
                torch.manual_seed(100 + t)
                single(t, fr, seq.intrinsics)
        n = single.n
        # union of the shards' edges = the single-GPU graph
        cnt = torch.tensor([sh.ii.numel()], device=dev)
        dist.all_reduce(cnt)
        a, b = single.poses_[:n].double(), sh.poses_[:n].double()
        ext = max(float((a[:, :3] - a[:1, :3]).norm(dim=-1).max()), 1e-3)
        dt = float((a[:, :3] - b[:, :3]).norm(dim=-1).max()) / ext
        dd = float((single.patches_[:n, :, 2] - sh.patches_[:n, :, 2]).abs().max())
        gathered = [torch.zeros_like(sh.poses_[:n]) for _ in range(world)]
        dist.all_gather(gathered, sh.poses_[:n].contiguous())
        same = all(bool((g == gathered[0]).all()) for g in gathered)
        gd = [torch.zeros_like(sh.patches_[:n]) for _ in range(world)]
        dist.all_gather(gd, sh.patches_[:n].contiguous())
        same_d = all(bool((g == gd[0]).all()) for g in gd)
        out[rank] = (int(cnt.item()), int(single.ii.numel()), dt, dd, same, same_d, sh.collective_calls,
                     sh.collective_bytes)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
def test_sharded_vo_matches_single_gpu_vo():
    """Ramp_vo(world_size=2): same 14 frames on both ranks, graph sharded by source frame; poses and depths are
    bit-identical across ranks and track the single-GPU VO (different fp32 summation order of [S|y])."""
    import torch.multiprocessing as mp
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_vo_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        e_sum, e_single, dt, dd, same, same_d, calls, nbytes = out[r]
        print("[sharded VO rank %d] edges %d/%d  pose dt/extent vs single-GPU %.3e  depth max|d| %.3e  collectives %d "
              "(%d bytes)" % (r, e_sum, e_single, dt, dd, calls, nbytes))
        assert e_sum == e_single
        assert same and same_d
        assert dt < 5e-2
