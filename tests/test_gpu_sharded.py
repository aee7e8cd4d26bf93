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
