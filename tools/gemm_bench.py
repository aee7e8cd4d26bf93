"""rvo_up_linear vs torch (cuBLASLt) on the update operator's layer shapes (CUDA events, L2 flushed)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rampvo_b200 import _lib  # noqa: E402


def timeit(fn, flush, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 45312
    L = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shapes = ((384, 0), (384, 1), (768, 0)) if len(sys.argv) < 3 else ((int(sys.argv[2]), 0),)
    for N, relu in shapes:
        x = torch.randn(M, 384, device="cuda").half()
        w = (torch.randn(N, 384, device="cuda") / 20).half()
        b = torch.randn(N, device="cuda").half()
        y = torch.empty(M, N, dtype=torch.float16, device="cuda")
        ours = timeit(lambda: L.rvo_up_linear(_lib.ptr(x), 384, _lib.ptr(w), _lib.ptr(b), M, 384, N, relu,
                                              _lib.ptr(y), N, _lib.stream_ptr()), flush)
        if relu:
            lib = timeit(lambda: torch._addmm_activation(b, x, w.t()), flush)
        else:
            lib = timeit(lambda: torch.nn.functional.linear(x, w, b), flush)
        fl = 2.0 * M * 384 * N
        class _NoFlush:
            def zero_(self):
                pass
        ours_w = timeit(lambda: L.rvo_up_linear(_lib.ptr(x), 384, _lib.ptr(w), _lib.ptr(b), M, 384, N, relu,
                                                _lib.ptr(y), N, _lib.stream_ptr()), _NoFlush())
        lib_w = timeit((lambda: torch._addmm_activation(b, x, w.t())) if relu else
                       (lambda: torch.nn.functional.linear(x, w, b)), _NoFlush())
        print(json.dumps({"M": M, "N": N, "relu": relu, "ours_us": round(ours, 2), "cublas_us": round(lib, 2),
                          "ours_warmL2_us": round(ours_w, 2), "cublas_warmL2_us": round(lib_w, 2),
                          "ours_TFLOPs": round(fl / ours / 1e6, 1), "cublas_TFLOPs": round(fl / lib / 1e6, 1)}))


def trace():
    import ctypes
    import numpy as np
    L = ctypes.CDLL(_lib.LIB_PATH)
    buf = (ctypes.c_longlong * (8 * 64))()
    assert L.rvo_up_trace(buf) == 0
    a = np.array(buf[:], dtype=np.int64).reshape(8, 64)
    t0 = a[0, 0]
    print("kernel end %d cycles after the W loads were issued" % (a[0, 2] - t0))
    print("tile  M.first_ready  M.last_issued   E.start    E.end    E.dur")
    for i in range(6):
        print("%3d %14d %14d %9d %8d %8d" % (i, a[2, i] - t0, a[3, i] - t0, a[4, i] - t0, a[5, i] - t0, a[5, i] - a[4, i]))
    print("producer: xempty seen at", [int(v - t0) for v in a[1, :32]])


if __name__ == "__main__":
    main()
    if os.environ.get("RVO_UP_TRACE"):
        trace()
