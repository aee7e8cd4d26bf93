#!/usr/bin/env python
"""train.py — drop-in for the reference's training loop (train.py:67-220) on synthetic TartanEvent-shaped batches
(BASELINE.json configs[4]): VONet.forward unroll (18 steps) -> flow + pose losses (train.py:29-64) -> AdamW +
OneCycleLR, gradient clipping, checkpoints with the reference's keys; data-parallel over the GPUs of one box with
torch.distributed / NCCL (one process per GPU, batch 1 per rank as ramp/utils.py:238 enforces), encoder in bf16.

    python train.py --steps 20                                             # one GPU
    python -m torch.distributed.run --nproc-per-node 8 train.py --steps 20  # 8 x B200, DDP gradient all-reduce

There is no dataset in this environment (scripts/download_tartanevent.sh needs the network): batches are generated
with the shapes of TartanEvent.__getitem__ (ramp/data_readers/TartanEvent.py:357-364).  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRAIN_CFG = {"input_mode": "MultiScale", "event_bias": True, "num_event_bins": 5, "n_frames": 15, "lr": 0.00008,
             "steps": 400000, "clip": 0.1, "pose_weight": 10.0, "flow_weight": 0.1, "pct_start": 0.01,
             "weight_decay": 1e-6, "batch_size": 1}              # config_net/MultiScale_TartanEvent.json:9-35


def kabsch_umeyama(A, B):
    """scale of the similarity aligning B to A (ramp/utils.py:389-399)"""
    EA, EB = A.mean(dim=0), B.mean(dim=0)
    var_a = ((A - EA).norm(dim=1) ** 2).mean()
    H = ((A - EA).T @ (B - EB)) / A.shape[0]
    D = torch.linalg.svdvals(H)
    return var_a / D.sum()


def compute_losses(traj, so, cfg, patch_size):
    """train.py:29-64: per unroll step, the best-pixel flow error of the valid patches and (from the third step
    on) the translation / rotation error of every relative pose after a scale alignment"""
    loss = 0.0
    for i, (v, x, y, P1, P2) in enumerate(traj):
        e = (x - y).norm(dim=-1)
        e = e.reshape(-1, patch_size ** 2)[(v > 0.5).reshape(-1)].min(dim=-1).values
        N = P1.shape[1]
        ii, jj = torch.meshgrid(torch.arange(N, device=x.device), torch.arange(N, device=x.device), indexing="ij")
        k = ii != jj
        ii, jj = ii[k], jj[k]
        P1, P2 = P1.inv(), P2.inv()
        t1, t2 = P1.matrix()[..., :3, 3], P2.matrix()[..., :3, 3]
        s = kabsch_umeyama(t2[0], t1[0]).detach().clamp(max=10.0)
        P1 = P1.scale(s.view(1, 1))
        dP = P1[:, ii].inv() * P1[:, jj]
        dG = P2[:, ii].inv() * P2[:, jj]
        e1 = (dP * dG.inv()).log()
        tr, ro = e1[..., 0:3].norm(dim=-1), e1[..., 3:6].norm(dim=-1)
        loss = loss + cfg["flow_weight"] * e.mean()
        if not so and i >= 2:
            loss = loss + cfg["pose_weight"] * (tr.mean() + ro.mean())
    return loss, e, ro, tr


def synthetic_batch(step, rank, n_frames, ht, wd, device):
    """events [1,T,5,H,W] (integer-valued stacks), images [1,T,3,H,W], poses [1,T,7], disps [1,T,H,W], K [1,T,4],
    mask [1,T] — one event stack per image (n_events_in_between = 1)"""
    g = torch.Generator(device=device).manual_seed(1000 * step + rank)
    T = n_frames
    ev = torch.poisson(torch.full((1, T, 5, ht, wd), 0.12, device=device), generator=g)
    ev = ev * (torch.randint(0, 2, ev.shape, generator=g, device=device) * 2 - 1)
    im = torch.rand(1, T, 3, ht, wd, generator=g, device=device) * 2 - 0.5
    poses = torch.zeros(1, T, 7, device=device)
    poses[..., 6] = 1.0
    poses[0, :, :3] = torch.cumsum(torch.randn(T, 3, generator=g, device=device) * 0.02, dim=0)
    disps = torch.rand(1, T, ht, wd, generator=g, device=device) * 0.8 + 0.2
    K = torch.tensor([wd / 2.0, wd / 2.0, wd / 2.0, ht / 2.0], device=device).repeat(1, T, 1)
    return ev, im, poses, disps, K, torch.ones(1, T, dtype=torch.bool)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--unroll", type=int, default=18)            # STEPS=18, train.py:151
    ap.add_argument("--ht", type=int, default=480)
    ap.add_argument("--wd", type=int, default=640)
    ap.add_argument("--frames", type=int, default=TRAIN_CFG["n_frames"])
    ap.add_argument("--ckpt", default=None)
    ap.add_argument("--save", default=None)
    ap.add_argument("--fp32-encoder", action="store_true")
    args = ap.parse_args()

    import torch.distributed as dist
    from rampvo_b200.lietorch import SE3
    from rampvo_b200.net import VONet
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(1234)
    net = VONet(TRAIN_CFG).to(dev).train()
    net.patchify.encoder_autocast = None if args.fp32_encoder else torch.bfloat16
    opt = torch.optim.AdamW(net.parameters(), lr=TRAIN_CFG["lr"], weight_decay=TRAIN_CFG["weight_decay"])
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=TRAIN_CFG["lr"], total_steps=TRAIN_CFG["steps"],
                                                pct_start=TRAIN_CFG["pct_start"], cycle_momentum=False,
                                                anneal_strategy="linear")
    step = 0
    if args.ckpt:                                                # train.py:93-106
        ck = torch.load(args.ckpt, map_location="cpu")
        step = ck["total_idx"]
        opt.load_state_dict(ck["optimizer_state_dict"])
        sched.load_state_dict(ck["scheduler_state_dict"])
        net.load_state_dict(OrderedDict((k.replace("module.", ""), v) for k, v in ck["model_state_dict"].items()),
                            strict=False)
    model = net
    if world > 1:       # the dead layer2 / conv2 weights (extractor.py:276-277) never receive a gradient
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], find_unused_parameters=True)

    def one_step(s):
        ev, im, poses, disps, K, mask = synthetic_batch(s, rank, args.frames, args.ht, args.wd, dev)
        opt.zero_grad()
        so = s < 1000 and args.ckpt is None                      # fix_repr_pose, train.py:148
        traj = model((ev, im, mask), SE3(poses).inv(), disps, K, STEPS=args.unroll, structure_only=so)
        loss, e, ro, tr = compute_losses(traj, so, TRAIN_CFG, net.P)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), TRAIN_CFG["clip"])
        opt.step()
        sched.step()
        return loss

    for s in range(args.warmup):
        one_step(step + s)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss = None
    for s in range(args.warmup, args.warmup + args.steps):
        loss = one_step(step + s)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    lv = float(loss.item())
    if args.save and rank == 0:                                  # train.py:180-196
        torch.save({"batch_idx": step + args.warmup + args.steps, "total_idx": step + args.warmup + args.steps,
                    "epoch": 0, "model_state_dict": model.state_dict(), "optimizer_state_dict": opt.state_dict(),
                    "scheduler_state_dict": sched.state_dict()}, args.save)
    if rank == 0:
        print(json.dumps({"metric": "train_clips_per_sec", "value": world * args.steps / dt, "unit": "clips/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * dt / args.steps, "loss": lv, "scaling": "weak", "data": "synthetic",
                          "config": {"workload": "train.py MultiScale, %d-frame clips %dx%d, unroll %d, batch 1 per "
                                                 "rank" % (args.frames, args.wd, args.ht, args.unroll),
                                     "encoder_dtype": "fp32" if args.fp32_encoder else "bf16",
                                     "parallelism": "DDP over %d ranks (NCCL gradient all-reduce)" % world
                                     if world > 1 else "single GPU"}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
