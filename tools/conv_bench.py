"""Per-layer timing of the tcgen05 implicit-GEMM convolution (rvo_conv2d_nhwc) against cuDNN on the encoder's layer
shapes, each launched alone (CUDA events, 50 launches back to back, L2 warm).  Usage: python tools/conv_bench.py"""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rampvo_b200.extractor import CL, _conv_tc  # noqa: E402

SHAPES = [("conv1 7x7s2 16->32 @480x640", 16, 0, 32, 7, 2, 3, 480, 640),
          ("layer1 3x3 32->32 @240x320", 32, 0, 32, 3, 1, 1, 240, 320),
          ("layer3.0.conv1 3x3s2 32+32->64", 32, 32, 64, 3, 2, 1, 240, 320),
          ("layer3.0.down 1x1s2 32+32->64", 32, 32, 64, 1, 2, 0, 240, 320),
          ("layer3 3x3 64->64 @120x160", 64, 0, 64, 3, 1, 1, 120, 160),
          ("conv3 1x1 64+64->128", 64, 64, 128, 1, 1, 0, 120, 160),
          ("conv3 1x1 64+64->384", 64, 64, 384, 1, 1, 0, 120, 160)]


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def trace():
    """-DRVO_DEBUG library only: globaltimer stamps of one CTA (ns relative to its start)"""
    import ctypes
    from rampvo_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    buf = (ctypes.c_ulonglong * 512)()
    for name, C0, C1, Cout, ks, sd, pd, H, W in SHAPES:
        conv = nn.Conv2d(C0 + C1, Cout, ks, stride=sd, padding=pd).cuda()
        x = torch.randn(1, C0, H, W, device="cuda").half().contiguous(memory_format=CL)
        x2 = torch.randn(1, C1, H, W, device="cuda").half().contiguous(memory_format=CL) if C1 else None
        for cta in (0, 100):
            L.rvo_conv_trace(buf, cta)
            for _ in range(3):
                _conv_tc(conv, x, x2, stats=True)
            L.rvo_conv_trace(buf, cta)
            t = [[buf[s * 64 + i] for i in range(64)] for s in range(8)]
            t0 = t[0][0]
            rel = lambda v: (v - t0) if v >= t0 else -1
            print("%s  CTA %d: setup %d  end %d ns" % (name, cta, rel(t[0][1]), rel(t[0][2])))
            for lt in range(6):
                if t[3][lt] < t0:
                    break
                print("    tile %d: acc/weights ready %6d  first A %6d  last MMA issued %6d  epilogue %6d .. %6d"
                      % (lt, rel(t[1][lt]), rel(t[2][lt]), rel(t[3][lt]), rel(t[4][lt]), rel(t[5][lt])))
            print("    producer 0: stage free / issued per K block: " +
                  " ".join("%d/%d" % (rel(t[6][i]), rel(t[7][i])) for i in range(16) if t[7][i] >= t0))


def one(idx):
    """a handful of launches of one layer shape (for `ncu -k regex:conv_tc_kernel`)"""
    name, C0, C1, Cout, ks, sd, pd, H, W = SHAPES[idx]
    conv = nn.Conv2d(C0 + C1, Cout, ks, stride=sd, padding=pd).cuda()
    x = torch.randn(1, C0, H, W, device="cuda").half().contiguous(memory_format=CL)
    x2 = torch.randn(1, C1, H, W, device="cuda").half().contiguous(memory_format=CL) if C1 else None
    for _ in range(6):
        _conv_tc(conv, x, x2, stats=True)
    torch.cuda.synchronize()
    print(name)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "trace":
        return trace()
    if len(sys.argv) > 2 and sys.argv[1] == "one":
        return one(int(sys.argv[2]))
    torch.backends.cudnn.benchmark = True
    lines = []
    for name, C0, C1, Cout, ks, sd, pd, H, W in SHAPES:
        conv = nn.Conv2d(C0 + C1, Cout, ks, stride=sd, padding=pd).cuda()
        x = torch.randn(1, C0, H, W, device="cuda").half().contiguous(memory_format=CL)
        x2 = torch.randn(1, C1, H, W, device="cuda").half().contiguous(memory_format=CL) if C1 else None
        xc = x if x2 is None else torch.cat((x, x2), 1).contiguous(memory_format=CL)
        w16 = conv.weight.detach().half().contiguous(memory_format=CL)
        b16 = conv.bias.detach().half()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            _conv_tc(conv, x, x2, stats=True)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(10):
                _conv_tc(conv, x, x2, stats=True)
        t_graph = timeit(g.replay, 20) / 10
        t_tc = timeit(lambda: _conv_tc(conv, x, x2, stats=True))
        t_tc_ns = timeit(lambda: _conv_tc(conv, x, x2, stats=False))
        t_dnn = timeit(lambda: F.conv2d(xc, w16, b16, stride=sd, padding=pd))
        flops = 2.0 * ((H + 2 * pd - ks) // sd + 1) * ((W + 2 * pd - ks) // sd + 1) * Cout * ks * ks * (C0 + C1)
        lines.append("%-36s tcgen05 %7.1f us (in a graph %6.1f, no stats %6.1f)   cuDNN %7.1f us   %.2f GFLOP"
                     % (name, t_tc, t_graph, t_tc_ns, t_dnn, flops / 1e9))
    out = "\n".join(lines)
    print(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "conv_bench.txt"), "w").write(out + "\n")


if __name__ == "__main__":
    main()
