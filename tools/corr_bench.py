"""Per-kernel timing of altcorr.corr_tiles on the default.yaml steady-state graph (torch.profiler / CUPTI).
Usage: python tools/corr_bench.py [default|precise]  (env RVO_CORR_DBG / RVO_CORR_LEGACY select variants)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_ops as O  # noqa: E402  (bench-side input generation only)
from rampvo_b200 import altcorr, synth  # noqa: E402


def main():
    cfg = "precise" if "precise" in sys.argv else "default"
    prob = synth.make_problem(cfg, {"default": 40, "precise": 80}[cfg], seed=0)
    M, E = prob["M"], prob["E"]
    gmap, pyr = synth.make_features(32, M * 32, seed=0)
    c = O.reproject(prob["poses"], prob["patches"], prob["intrinsics"], prob["ii"], prob["jj"], prob["kk"]).astype(np.float32)
    g_t = torch.from_numpy(gmap).cuda().permute(0, 3, 1, 2)[None]
    p_t = [torch.from_numpy(p).cuda().permute(0, 3, 1, 2)[None] for p in pyr]
    c_t = torch.from_numpy(c).cuda()[None]
    k_t, j_t = torch.from_numpy(prob["kk"]).cuda(), torch.from_numpy(prob["jj"]).cuda()
    out = torch.zeros(1, E, 1008, dtype=torch.float16, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush_r = torch.zeros(64 << 20, dtype=torch.int32, device="cuda")     # read after the write: clean lines in the L2
    for _ in range(3):
        altcorr.corr_tiles(g_t, p_t, c_t, k_t, j_t, M * 32, 32, out=out)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    n = 10
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            flush.zero_()
            if "dirty" not in sys.argv:
                flush_r.max()
            altcorr.corr_tiles(g_t, p_t, c_t, k_t, j_t, M * 32, 32, out=out)
        torch.cuda.synchronize()
    agg = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and "rvo::" in e.name:
            a = agg.setdefault(e.name.split("(")[0], [0, 0.0])
            a[0] += 1
            a[1] += e.device_time
    res = {k: round(v[1] / v[0], 2) for k, v in agg.items()}
    res["total_us"] = round(sum(v[1] for v in agg.values()) / n, 2)
    res["dbg"] = os.environ.get("RVO_CORR_DBG", "0")
    print(json.dumps(res))
    if "trace" in sys.argv:
        trace()




def trace():
    """RVO_CORR_DBG=16 python tools/corr_bench.py trace: per-block clock stamps of CTA 0"""
    import ctypes
    from rampvo_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    buf = (ctypes.c_longlong * (8 * 256))()
    assert L.rvo_corr_trace(buf) == 0
    a = np.array(buf[:], dtype=np.int64).reshape(8, 256)
    t0 = a[2, 0]
    names = ["A.aempty", "A.publ", "M.top", "M.ready", "M.issued", "E.tfull", "E.done"]
    print("blk " + " ".join("%9s" % n for n in names) + "   issue  M.period E.dur")
    for i in range(24, 64):
        row = [int(a[s, i] - t0) for s in range(7)]
        print("%3d " % i + " ".join("%9d" % v for v in row) + "  %6d %8d %5d" % (row[4] - row[3], int(a[4, i] - a[4, i - 1]), row[6] - row[5]))
    per = np.diff(a[4, 8:200]).mean()
    print("mean cycles/block (MMA issue to issue): %.0f; mean E.dur %.0f; mean issue %.0f; mean M wait %.0f" % (
        per, (a[6, 8:200] - a[5, 8:200]).mean(), (a[4, 8:200] - a[3, 8:200]).mean(), (a[3, 8:200] - a[2, 8:200]).mean()))


if __name__ == "__main__":
    main()
