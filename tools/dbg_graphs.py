import sys, os, time, torch
sys.path.insert(0, '/root/repo')
import bench
from rampvo_b200 import synth
dev = torch.device("cuda", 0)
seq = synth.SyntheticSequence(seed=0, device=dev)
frames = [seq.frame(t) for t in range(bench.SETUP_FRAMES + 60)]
vo = bench.build_vo(dev)
with torch.no_grad():
    for t in range(bench.SETUP_FRAMES + 60):
        k0 = set(vo._ugraphs.keys())
        t0 = time.perf_counter()
        vo(t, frames[t], seq.intrinsics)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        new = set(vo._ugraphs.keys()) - k0
        if t >= bench.SETUP_FRAMES - 3 and (new or dt > 3.0):
            print(t, "n", vo.n, "E", vo.ii.numel(), "ms %.1f" % dt, "new graph", [(k[0], k[1], k[2]) for k in new], "cached", len(vo._ugraphs))
print("final keys", [(k[0], k[1], k[2]) for k in vo._ugraphs])
