"""altcorr lookup timed on the patch graph of the RUNNING visual odometry (the state bench.py's roofline block
measures: coords of a real steady-state update, rows clustered where the patch selection put them), next to the tile
statistics that decide the kernel's block structure.  tools/corr_bench.py times the same call on a synthetic
uniformly spread problem.  Usage: python tools/corr_bench_vo.py [frames]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rampvo_b200 import synth  # noqa: E402


def tile_stats(coords, jj, scale, W, H, step=9, win=8, r=3):
    """rows per 16x16 tile (tiles step by 9) and 128-row blocks, as rvo_corr_tiles bins them"""
    x0 = torch.floor(coords[0, :, 0].reshape(-1, 9) * scale).long() - r
    y0 = torch.floor(coords[0, :, 1].reshape(-1, 9) * scale).long() - r
    ok = (x0 > -win) & (x0 < W) & (y0 > -win) & (y0 < H)
    tx, ty = (x0 + win) // step, (y0 + win) // step
    key = (jj[:, None] * 64 + ty) * 64 + tx
    u, c = torch.unique(key[ok], return_counts=True)
    blocks = ((c + 127) // 128)
    return {"rows": int(ok.sum()), "tiles": int(u.numel()), "rows_per_tile_mean": float(c.float().mean()),
            "rows_per_tile_p50": int(c.median()), "rows_per_tile_max": int(c.max()), "blocks": int(blocks.sum()),
            "blocks_le64": int((c <= 64).sum()), "tiles_multi_block": int((c > 128).sum())}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    dev = torch.device("cuda", 0)
    seq = synth.SyntheticSequence(seed=0, device=dev)
    with torch.no_grad():
        vo = bench.build_vo(dev)
        for t in range(n):
            vo(t, seq.frame(t), seq.intrinsics)
        vo.sync()
        torch.cuda.synchronize()
        coords = vo.reproject()
        E = int(vo.ii.numel())
        st = {"E": E,
              "level1": tile_stats(coords, vo.jj % vo.mem, 1.0, vo.wd // 4, vo.ht // 4),
              "level2": tile_stats(coords, vo.jj % vo.mem, 0.25, vo.wd // 16, vo.ht // 16)}
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        flush_r = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
        vo.corr_tiles(coords)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()          # the call as the frame runs it (eager, the host is slower than the GPU)
        with torch.cuda.graph(g):
            vo.corr_tiles(coords)
        g.replay()
        from torch.profiler import ProfilerActivity, profile
        ts = []
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10):
                flush.zero_()
                flush_r.max()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                g.replay()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
        agg = {}
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA and "rvo::" in e.name:
                v = agg.setdefault(e.name.split("(")[0], [0, 0.0])
                v[0] += 1
                v[1] += e.device_time
        st["kernels_us"] = {k: round(v[1] / v[0], 2) for k, v in agg.items()}
        st["graph_replay_us_events"] = round(sum(ts) / len(ts), 2)
    print(json.dumps(st))


if __name__ == "__main__":
    main()
