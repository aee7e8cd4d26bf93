"""Host-side wall time per section of Ramp_vo.__call__ at steady state (no device syncs added except
at the frame end).  Usage: python tools/host_profile.py
Sharded mode: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/host_profile.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rampvo_b200 import synth  # noqa: E402
from rampvo_b200.Ramp_vo import Ramp_vo  # noqa: E402

T = {}


def wrap(cls, name):
    f = getattr(cls, name)

    def g(self, *a, **k):
        t = time.perf_counter()
        r = f(self, *a, **k)
        T[name] = T.get(name, 0.0) + time.perf_counter() - t
        return r
    setattr(cls, name, g)


def main():
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    seq = synth.SyntheticSequence(seed=0, device=dev)
    n = 30
    frames = [seq.frame(t) for t in range(bench.SETUP_FRAMES + n + 3)]
    vo = bench.build_vo(dev, world_size=world, rank=rank)
    for t in range(bench.SETUP_FRAMES + 3):
        vo(t, frames[t], seq.intrinsics)
    for name in ("__call__", "sync", "_keyframe_defer", "update", "keyframe", "append_factors", "remove_factors",
                 "_update_graphed", "_edges_forw", "_edges_back", "_edges_step", "_keyframe_finish", "_pair_counts",
                 "_update_sharded", "_keyframe_begin", "_graph_plans"):
        wrap(Ramp_vo, name)
    if world > 1:
        from rampvo_b200 import sharded as sh

        def wrapf(modl, name):
            f = getattr(modl, name)

            def g(*a, **k):
                t = time.perf_counter()
                r = f(*a, **k)
                T[name] = T.get(name, 0.0) + time.perf_counter() - t
                return r
            setattr(modl, name, g)
        wrapf(sh, "sharded_BA_fused")
        wrapf(sh, "exchange_depths_owned")
    from rampvo_b200 import Ramp_vo as mod
    wrap(mod._PatchifyGraph, "run")
    wrap(mod._UpdateGraph, "run")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(bench.SETUP_FRAMES + 3, bench.SETUP_FRAMES + 3 + n):
        vo(t, frames[t], seq.intrinsics)
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    if rank == 0:
        print("frame wall %.3f ms (world %d)" % (tot / n * 1e3, world))
        for k, v in sorted(T.items(), key=lambda x: -x[1]):
            print("  %-22s %.3f ms/frame" % (k, v / n * 1e3))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
