"""Per-stream occupancy of steady-state frames (torch.profiler / CUPTI): for every CUDA stream the busy time per
frame, the union over streams (time with at least one kernel running) and the overlap between the encoder stream
and the update stream.  Usage: [RVO_SM_SPLIT=enc,upd] python tools/stream_timeline.py [n_frames]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rampvo_b200 import synth  # noqa: E402


def union(iv):
    iv = sorted(iv)
    tot, cs, ce = 0, None, None
    for s, e in iv:
        if cs is None:
            cs, ce = s, e
        elif s <= ce:
            ce = max(ce, e)
        else:
            tot += ce - cs
            cs, ce = s, e
    return tot + (ce - cs if cs is not None else 0)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda", 0)
    seq = synth.SyntheticSequence(seed=0, device=dev)
    frames = [seq.frame(t) for t in range(bench.SETUP_FRAMES + n + 3)]
    with torch.no_grad():
        vo = bench.build_vo(dev)
        for t in range(bench.SETUP_FRAMES + 3):
            vo(t, frames[t], seq.intrinsics)
        torch.cuda.synchronize()
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for t in range(bench.SETUP_FRAMES + 3, bench.SETUP_FRAMES + 3 + n):
                vo(t, frames[t], seq.intrinsics)
            torch.cuda.synchronize()
    per = {}
    for e in prof.profiler.kineto_results.events():
        if e.device_type() != torch.autograd.DeviceType.CUDA or e.duration_ns() <= 0:
            continue
        per.setdefault(e.device_resource_id(), []).append((e.start_ns(), e.start_ns() + e.duration_ns(), e.name()))
    t0 = min(s for v in per.values() for s, _, _ in v)
    t1 = max(e for v in per.values() for _, e, _ in v)
    span = (t1 - t0) / n / 1e3
    lines = ["frames=%d  SM_SPLIT=%s  wall/frame=%.1f us  any-kernel-running/frame=%.1f us" % (
        n, os.environ.get("RVO_SM_SPLIT", "-"), span, union([(s, e) for v in per.values() for s, e, _ in v]) / n / 1e3)]
    for sid, v in sorted(per.items(), key=lambda kv: -sum(e - s for s, e, _ in kv[1])):
        busy = union([(s, e) for s, e, _ in v]) / n / 1e3
        top = {}
        for s, e, nm in v:
            top[nm[:40]] = top.get(nm[:40], 0) + (e - s)
        names = ", ".join("%s %.0f" % (k, t / n / 1e3) for k, t in sorted(top.items(), key=lambda kv: -kv[1])[:4])
        lines.append("stream %-4d kernels/frame=%6.1f busy/frame=%8.1f us   [%s]" % (sid, len(v) / n, busy, names))
    out = "\n".join(lines)
    print(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "stream_timeline.txt"), "a").write(out + "\n\n")


if __name__ == "__main__":
    main()
