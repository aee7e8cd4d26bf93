"""Per-stretch CUDA-event times of the update operator on chains (rvo_up_chain) on a default.yaml-sized graph, next
to its layer-by-layer form, plus single-epilogue micro chains that isolate where a tile's time goes.
Usage: python tools/chain_bench.py [n_frames] [--quick]   (under ncu: -k regex:up_chain)"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rampvo_b200 import _lib, net as N  # noqa: E402


def graph(n_frames, M, seed=0):
    rng = np.random.RandomState(seed)
    ii, jj, kk = [], [], []
    for i in range(n_frames):
        for p in range(M):
            for j in range(max(0, i - 6), min(n_frames, i + 7)):
                ii.append(i); jj.append(j); kk.append(i * M + p)
    perm = rng.permutation(len(ii))
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.int64)[perm]).cuda()
    return t(ii), t(jj), t(kk)


def micro_only():
    """python tools/chain_bench.py --micro: plain Linear+ReLU chains at row counts that put 1 or 3 tiles on 74 / 148
    CTAs — does a layer's time follow the number of CTAs streaming weights from L2, or the tiles per CTA?"""
    g = torch.Generator(device="cuda").manual_seed(5)
    w = (torch.randn(384, 384, generator=g, device="cuda") / 20).half()
    b = torch.randn(384, generator=g, device="cuda").half()
    st = _lib.stream_ptr("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    import ctypes as C
    L = _lib.lib()
    for cs, ctas, tiles in ((1, 37, 3), (1, 74, 3), (1, 148, 3), (2, 148, 3), (4, 148, 3), (4, 148, 6), (0, 148, 3)):
        _lib.check(L.rvo_up_chain_set_cluster(cs), "set_cluster")
        got_cs, got_ctas = C.c_int(), C.c_int()
        _lib.check(L.rvo_up_chain_info(C.byref(got_cs), C.byref(got_ctas)), "info")
        E = ctas * tiles * 128
        x16 = torch.randn(E, 384, generator=g, device="cuda").half()
        o16 = torch.empty(E, 384, device="cuda", dtype=torch.float16)
        res = {}
        for nl in (1, 5):
            fn = lambda: N.run_chain(E, _lib.PRO_ROWS, [(w, b, _lib.EPI_RELU, {})] * (nl - 1) +
                                     [(w, b, _lib.EPI_STORE16, {"y16": o16, "ldy": 384})], st, a16=x16, lda=384)
            fn()
            ts = []
            for _ in range(7):
                flush.zero_()
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); c.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(c) * 1e3)
            res[nl] = sorted(ts)[3]
        print(json.dumps({"cluster": got_cs.value, "wave_ctas": got_ctas.value, "ctas": ctas, "tiles_per_cta": tiles, "E": E, "store16_only_us": round(res[1], 1),
                          "relu_x4_store16_us": round(res[5], 1),
                          "us_per_relu_layer_per_tile": round((res[5] - res[1]) / 4 / tiles, 2)}))


def trace():
    """python tools/chain_bench.py --trace (library built with `python -m rampvo_b200.build --debug`): per-layer
    timeline of CTA 0 for Linear+ReLU x4 + STORE16 on 3 tiles per CTA"""
    import ctypes as C
    L = C.CDLL(_lib.LIB_PATH)
    g = torch.Generator(device="cuda").manual_seed(5)
    w = (torch.randn(384, 384, generator=g, device="cuda") / 20).half()
    b = torch.randn(384, generator=g, device="cuda").half()
    E = 148 * 3 * 128
    x16 = torch.randn(E, 384, generator=g, device="cuda").half()
    o16 = torch.empty(E, 384, device="cuda", dtype=torch.float16)
    fn = lambda: N.run_chain(E, _lib.PRO_ROWS, [(w, b, _lib.EPI_RELU, {})] * 4 + [(w, b, _lib.EPI_STORE16, {"y16": o16, "ldy": 384})],
                             _lib.stream_ptr("cuda"), a16=x16, lda=384)
    fn(); fn()
    buf = (C.c_longlong * (4 * 64))()
    assert L.rvo_up_chain_trace(buf) == 0
    t = np.array(buf[:]).reshape(4, 64)
    t0 = t[0, 0]
    print("layer: operands_ready  last_mma_issued  accumulator_ready  epilogue_done   (cycles since first)")
    for i in range(15):
        print("%2d: %8d %8d %8d %8d   mma %6d  epi %6d  gap->next %6d" % (
            i, t[0, i] - t0, t[1, i] - t0, t[2, i] - t0, t[3, i] - t0, t[2, i] - t[0, i], t[3, i] - t[2, i],
            (t[0, i + 1] - t[3, i]) if i < 14 else 0))


def main():
    if "--micro" in sys.argv:
        return micro_only()
    if "--trace" in sys.argv:
        return trace()
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 40
    quick = "--quick" in sys.argv
    torch.manual_seed(0)
    up = N.Update(3).cuda().eval()
    ii, jj, kk = graph(n_frames, 96)
    E = ii.numel()
    g = torch.Generator(device="cuda").manual_seed(5)
    net = torch.randn(1, E, 384, generator=g, device="cuda")
    imap = torch.randn(n_frames * 96, 384, generator=g, device="cuda").half()
    corr = torch.randn(1, E, 1008, generator=g, device="cuda").half()
    plans = N.GraphPlans(ii, jj, kk)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    orig = N.run_chain

    def timed(M, prologue, layers, stream, **kw):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig(M, prologue, layers, stream, **kw)
        b.record()
        times_order.append((len(layers), a, b))

    reps = 1 if quick else 5
    acc = {}
    for a in sys.argv:
        if a.startswith("--cluster="):
            _lib.check(_lib.lib().rvo_up_chain_set_cluster(int(a.split("=")[1])), "set_cluster")
    with torch.no_grad():
        up._forward_chains(net, (imap, kk, 0), corr, ii, jj, kk, plans)       # warm-up
        torch.cuda.synchronize()
        N.run_chain = timed
        for _ in range(reps):
            times_order = []
            flush.zero_()
            up._forward_chains(net, (imap, kk, 0), corr, ii, jj, kk, plans)
            torch.cuda.synchronize()
            for i, (nl, a, b) in enumerate(times_order):
                acc.setdefault((i, nl), []).append(a.elapsed_time(b) * 1e3)
        N.run_chain = orig

        def whole(fn):
            ts = []
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            return sorted(ts)[len(ts) // 2]
        t_chain = whole(lambda: up._forward_chains(net, (imap, kk, 0), corr, ii, jj, kk, plans))
        t_layer = None if quick else whole(lambda: up._forward_fused(net, (imap, kk, 0), corr, ii, jj, kk, plans))
    out = {"E": E, "update_on_chains_us": round(t_chain, 1), "update_layered_us": t_layer and round(t_layer, 1),
           "stretches_us": {"stretch%d_%dlayers" % (i + 1, nl): round(sorted(v)[len(v) // 2], 1) for (i, nl), v in sorted(acc.items())}}
    print(json.dumps(out))
    if quick:
        return
    # micro chains: one epilogue kind at a time on the same E
    x16 = torch.randn(E, 384, generator=g, device="cuda").half()
    x32 = N.chain_block32(torch.randn(E, 384, generator=g, device="cuda"))
    o32 = torch.empty(N.chain_rows(E), 384, device="cuda")
    o16 = torch.empty(E, 384, device="cuda", dtype=torch.float16)
    w = (torch.randn(384, 384, generator=g, device="cuda") / 20).half()
    b = torch.randn(384, generator=g, device="cuda").half()
    st = _lib.stream_ptr("cuda")
    micro = {
        "1x STORE16": lambda: orig(E, _lib.PRO_ROWS, [(w, b, _lib.EPI_STORE16, {"y16": o16, "ldy": 384})], st, a16=x16, lda=384),
        "RELU + STORE16": lambda: orig(E, _lib.PRO_ROWS, [(w, b, _lib.EPI_RELU, {}), (w, b, _lib.EPI_STORE16, {"y16": o16, "ldy": 384})], st, a16=x16, lda=384),
        "RELU x4 + STORE16": lambda: orig(E, _lib.PRO_ROWS, [(w, b, _lib.EPI_RELU, {})] * 4 + [(w, b, _lib.EPI_STORE16, {"y16": o16, "ldy": 384})], st, a16=x16, lda=384),
        "RELU + RES(out32+out16)": lambda: orig(E, _lib.PRO_ROWS, [(w, b, _lib.EPI_RELU, {}), (w, b, _lib.EPI_RES, {})], st, a16=x16, lda=384, res32=x32, out32=o32, out16=o16),
    }
    for name, fn in micro.items():
        fn()
        print(json.dumps({"micro": name, "us": round(whole(fn), 1)}))


if __name__ == "__main__":
    main()
