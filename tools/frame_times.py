"""Per-frame GPU time of the device-resident stream right after the set-up frames (CUDA event after every call):
does the frame rate have a warm-up transient?  Usage: python tools/frame_times.py [n_frames]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rampvo_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    dev = torch.device("cuda", 0)
    seq = synth.SyntheticSequence(seed=0, device=dev)
    frames = [seq.frame(t) for t in range(bench.SETUP_FRAMES + n)]
    torch.cuda.synchronize()
    with torch.no_grad():
        vo = bench.build_vo(dev)
        for t in range(bench.SETUP_FRAMES):
            vo(t, frames[t], seq.intrinsics)
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for i in range(n):
            t = bench.SETUP_FRAMES + i
            vo(t, frames[t], seq.intrinsics)
            evs[i + 1].record()
        vo.sync()
        torch.cuda.synchronize()
    ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
    for i in range(0, n, 10):
        print("frames %3d-%3d: %s   mean %.2f ms  (n=%d E=%d graphs=%d)" % (
            i, i + 9, " ".join("%.2f" % v for v in ms[i:i + 10]), sum(ms[i:i + 10]) / len(ms[i:i + 10]), vo.n,
            vo.ii.numel(), len(vo._ugraphs)))


if __name__ == "__main__":
    main()
