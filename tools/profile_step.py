"""Per-kernel time breakdown of steady-state frames with torch.profiler (CUPTI), no ncu replay.
Usage: python tools/profile_step.py [n_frames]  -> prints a table and writes gpurun_out/step_profile.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rampvo_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    dev = torch.device("cuda", 0)
    seq = synth.SyntheticSequence(seed=0, device=dev)
    frames = [seq.frame(t) for t in range(bench.SETUP_FRAMES + n + 3)]
    with torch.no_grad():
        vo = bench.build_vo(dev)
        for t in range(bench.SETUP_FRAMES + 3):
            vo(t, frames[t], seq.intrinsics)
        torch.cuda.synchronize()
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for t in range(bench.SETUP_FRAMES + 3, bench.SETUP_FRAMES + 3 + n):
                vo(t, frames[t], seq.intrinsics)
            torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    agg = {}
    for e in evs:
        a = agg.setdefault(e.name[:90], [0, 0.0])
        a[0] += 1
        a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    tot = sum(v[1] for v in agg.values())
    lines = ["frames=%d  E=%d  kernels/frame=%.0f  gpu-busy us/frame=%.1f" % (n, vo.ii.numel(), len(evs) / n, tot / n)]
    for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:60]:
        lines.append("%-92s n/frame=%6.1f us/frame=%9.1f avg=%8.2f %5.1f%%" % (name, c / n, t / n, t / c, 100 * t / tot))
    out = "\n".join(lines)
    print(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "step_profile.txt"), "w").write(out + "\n")


if __name__ == "__main__":
    main()
