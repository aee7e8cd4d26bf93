"""Kernel-level timings of the hot-path ops on the synthetic steady-state graphs (CUDA events).
Usage: python tools/microbench.py [default|precise|fast] — prints one JSON line per op."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import ref_ops as O  # noqa: E402  (bench-side checker only)
from rampvo_b200 import altcorr, fastba, projective_ops as pops, synth  # noqa: E402
from rampvo_b200.lietorch import SE3  # noqa: E402


def timeit(fn, iters=20, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "default"
    nf = {"default": 40, "precise": 80, "fast": 40, "cfg1": 8}[cfg]
    prob = synth.make_problem(cfg, nf, seed=0)
    M = prob["M"]
    E = prob["E"]
    dev = "cuda"
    gmap, pyr = synth.make_features(32, M * 32, seed=0)
    t = {k: torch.from_numpy(prob[k]).to(dev) for k in ("ii", "jj", "kk")}
    poses = torch.from_numpy(prob["poses"]).to(dev)[None]
    patches = torch.from_numpy(prob["patches"]).to(dev)[None]
    intr = torch.from_numpy(prob["intrinsics"]).to(dev)[None]
    g_t = torch.from_numpy(gmap).to(dev).permute(0, 3, 1, 2)[None]
    p_t = [torch.from_numpy(p).to(dev).permute(0, 3, 1, 2)[None] for p in pyr]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    coords = pops.reproject_cf(SE3(poses), patches, intr, t["ii"], t["jj"], t["kk"])
    out = torch.empty(1, E, 882, dtype=torch.float16, device=dev)
    res = {"config": cfg, "E": E}

    med, mn = timeit(lambda: pops.reproject_cf(SE3(poses), patches, intr, t["ii"], t["jj"], t["kk"], out=coords), flush=flush)
    res["reproject_us"] = med
    med, mn = timeit(lambda: altcorr.corr_pyramid(g_t, p_t, coords, t["kk"], t["jj"], M * 32, 32, 3, out=out), flush=flush)
    # algorithmic bytes, SURVEY.md 8(d): output + coords/indices + unique gmap patches + unique feature pixels
    U = len(np.unique(prob["kk"]))
    F = len(np.unique(prob["jj"] % 32))
    s = 2
    byt = E * 882 * s + E * (18 * 4 + 16) + U * 128 * 9 * s
    for (h, w) in ((120, 160), (30, 40)):
        byt += min(F * 128 * h * w * s, E * 100 * 128 * s)
    res["corr_us"] = med
    res["corr_us_min"] = mn
    res["corr_alg_MB"] = byt / 1e6
    res["corr_GBs"] = byt / med / 1e3
    out_t = torch.zeros(1, E, 1008, dtype=torch.float16, device=dev)
    try:
        med, mn = timeit(lambda: altcorr.corr_tiles(g_t, p_t, coords, t["kk"], t["jj"], M * 32, 32, out=out_t), flush=flush)
        res["corr_tiles_us"] = med
        res["corr_tiles_GBs"] = byt / med / 1e3
    except Exception as ex:  # noqa: BLE001
        res["corr_tiles_error"] = str(ex)[:200]
    med, mn = timeit(lambda: fastba.neighbors(t["kk"], t["jj"], kmax=patches.shape[1], jmax=poses.shape[1]), flush=flush)
    res["neighbors_us"] = med
    tgt = (coords[0, :, :, 1, 1] + torch.from_numpy(prob["noise"]).to(dev))[None].contiguous()
    wgt = torch.from_numpy(prob["weight"]).to(dev)[None]
    lm = torch.tensor([1e-4], device=dev)
    p0, q0 = poses.clone(), patches.clone()

    def ba():
        poses.copy_(p0)
        patches.copy_(q0)
        fastba.BA(poses, patches, intr, tgt, wgt, lm, t["ii"], t["jj"], t["kk"], prob["t0"], prob["t1"], M, 2)
    med, mn = timeit(ba, flush=flush)
    res["ba2_us"] = med
    res["ba2_us_min"] = mn
    print(json.dumps(res))


if __name__ == "__main__":
    main()
