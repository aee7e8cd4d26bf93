"""Drop-in for ramp.net (ramp/net.py): VONet with `.patchify` (Patchifier) and `.update` (Update).

Parameter names match the reference so that its checkpoints load unchanged.  The forward passes are
restructured around the B200 kernels of librampvo_b200.so:

  Update.forward    neighbours come from the device-side graph plan (no CPU round trip,
                    ramp/fastba/ba.cpp:59-97), the two SoftAgg blocks are one segmented pass each
                    over pre-grouped edges (no torch.unique / torch_scatter, ramp/blocks.py:42-48);
                    the hidden state stays fp32 between stages like the reference under autocast
                    (LayerNorm returns fp32, SURVEY.md appendix "dtype drift").
  Patchifier        patch selection + the four altcorr.patchify gathers with the bilinear blend
                    fused; gmap is written straight into the caller's channels-last ring slot.
The 384-input Linear layers run on the hand-written tcgen05 GEMM (rvo_up_linear); only the first
correlation layer (K = 1008) still calls cuBLAS.
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

import numpy as np

from . import _lib, altcorr, fastba
from . import projective_ops as pops
from .ba import BA
from .extractor import MergerLSTMsceneEncoder, MultiScaleMergerDoubleNet
from .lietorch import SE3
from .vo_utils import coords_from_topk_events, flatmeshgrid, get_channel_dim, preprocess_input

DIM = 384


def _addr(t, off=0):
    """tensor (+ byte offset) -> ctypes pointer; None and ready-made pointers pass through"""
    if t is None or isinstance(t, ctypes.c_void_p):
        return t
    return ctypes.c_void_p(t.data_ptr() + off)


def chain_rows(M):
    """rows of a tile-blocked fp32 array that holds M edge rows (whole 128-row tiles)"""
    return (M + 127) // 128 * 128


def chain_block32(x):
    """[M, 384] fp32 row-major -> the tile-blocked layout rvo_up_chain keeps its private fp32 arrays in
    ([tile][16-byte chunk 0..95][row 0..127][4 floats], include/rampvo_b200.h); tests / debugging only"""
    M = x.shape[0]
    T = chain_rows(M) // 128
    xp = torch.zeros(T * 128, DIM, dtype=x.dtype, device=x.device)
    xp[:M] = x
    return xp.view(T, 128, DIM // 4, 4).permute(0, 2, 1, 3).contiguous().view(T * 128, DIM)


def chain_unblock32(xb, M):
    T = xb.shape[0] // 128
    return xb.view(T, DIM // 4, 128, 4).permute(0, 2, 1, 3).reshape(T * 128, DIM)[:M]


def run_chain(M, prologue, layers, stream, **kw):
    """rvo_up_chain (include/rampvo_b200.h): `layers` = [(w16, bias16, epilogue, {K, gamma, beta, y16, ldy})],
    keyword arguments = the other fields of rvo_chain_t (tensors, ctypes pointers or ints)."""
    c = _lib.Chain()
    c.M, c.n_layers, c.prologue = M, len(layers), prologue
    for k, v in kw.items():
        setattr(c, k, v if isinstance(v, int) else _addr(v))
    for i, (w, b, epi, extra) in enumerate(layers):
        ly = c.layer[i]
        extra = dict(extra)
        ly.w16, ly.bias16, ly.K, ly.epilogue = _addr(w), _addr(b), extra.pop("K", DIM), epi
        for k, v in extra.items():
            setattr(ly, k, v if isinstance(v, int) else _addr(v))
    _lib.check(_lib.lib().rvo_up_chain(ctypes.byref(c), stream), "rvo_up_chain")


class GatedResidual(nn.Module):
    """ramp/blocks.py:15-31: x + sigmoid(W_g x) * W_2 relu(W_1 x)"""

    def __init__(self, dim):
        super().__init__()
        self.gate = nn.Sequential(nn.Linear(dim, dim), nn.Sigmoid())
        self.res = nn.Sequential(nn.Linear(dim, dim), nn.ReLU(inplace=True), nn.Linear(dim, dim))

    def forward(self, x):
        return x + self.gate(x) * self.res(x)


class SoftAgg(nn.Module):
    """ramp/blocks.py:33-50 parameters; the aggregation itself is Update._soft_agg."""

    def __init__(self, dim=512, expand=True):
        super().__init__()
        self.dim = dim
        self.expand = expand
        self.f = nn.Linear(dim, dim)
        self.g = nn.Linear(dim, dim)
        self.h = nn.Linear(dim, dim)


class GraphPlans:
    """Device-side bookkeeping shared by one update: the (kk, jj) plan gives `neighbors` and the
    agg_kk groups, the (ii*12345+jj) plan gives the agg_ij groups (net.py:77,84-85)."""

    def __init__(self, ii, jj, kk, kmax=0, jmax=0, max_patches=0, max_pairs=0):
        L = _lib.lib()
        self.E = E = ii.numel()
        # upper bounds on the number of groups (0 = E); they size the per-group GEMM of SoftAgg.h
        self.cap_k = min(E, max_patches) if max_patches else E
        self.cap_ij = min(E, max_pairs) if max_pairs else E
        dev = ii.device
        nb = L.rvo_plan_bytes(E)
        self.plan_k = torch.empty(nb, dtype=torch.uint8, device=dev)
        self.plan_ij = torch.empty(nb, dtype=torch.uint8, device=dev)
        self.ix = torch.empty(E, dtype=torch.int64, device=dev)
        self.jx = torch.empty(E, dtype=torch.int64, device=dev)
        if E == 0:
            return
        st = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            _lib.check(L.rvo_graph_plan(_lib.ptr(kk), _lib.ptr(jj), E, kmax, jmax, _lib.ptr(self.plan_k),
                                        nb, st), "rvo_graph_plan")
            _lib.check(L.rvo_plan_neighbors(_lib.ptr(self.plan_k), E, _lib.ptr(self.ix),
                                            _lib.ptr(self.jx), st), "rvo_plan_neighbors")
            if jmax:
                # agg_ij groups by ii * 12345 + jj (net.py:85); any key that is injective on (ii, jj) and keeps
                # their lexicographic order gives the same groups in the same order: ii * 2^b + jj needs 2 radix
                # passes where the reference's multiplier needs 4 (and the second sort key is redundant here)
                J = 1 << max(int(jmax - 1).bit_length(), 1)
                key_ij = ii * J + jj
                _lib.check(L.rvo_graph_plan(_lib.ptr(key_ij), _lib.ptr(self._zeros(E, dev)), E, jmax * J, 1,
                                            _lib.ptr(self.plan_ij), nb, st), "rvo_graph_plan")
            else:
                key_ij = ii * 12345 + jj
                _lib.check(L.rvo_graph_plan(_lib.ptr(key_ij), _lib.ptr(jj), E, 0, 0, _lib.ptr(self.plan_ij), nb, st),
                           "rvo_graph_plan")

    _zero_cache = {}

    @classmethod
    def _zeros(cls, E, dev):
        """an all-zero int64 secondary key of at least E entries (cached per device: never written)"""
        z = cls._zero_cache.get(str(dev))
        if z is None or z.numel() < E:
            z = cls._zero_cache[str(dev)] = torch.zeros(max(E, 1 << 16), dtype=torch.int64, device=dev)
        return z

    def edge_groups(self):
        """device pointers (ctypes) to the group id of every edge in the kk plan and in the (ii, jj) plan"""
        L = _lib.lib()
        out = []
        for plan in (self.plan_k, self.plan_ij):
            p = ctypes.c_void_p()
            _lib.check(L.rvo_plan_edge_groups(_lib.ptr(plan), self.E, ctypes.byref(p)), "rvo_plan_edge_groups")
            out.append(p)
        return out


class Update(nn.Module):
    """ramp/net.py:34-90."""

    def __init__(self, p):
        super().__init__()
        self.c1 = nn.Sequential(nn.Linear(DIM, DIM), nn.ReLU(inplace=True), nn.Linear(DIM, DIM))
        self.c2 = nn.Sequential(nn.Linear(DIM, DIM), nn.ReLU(inplace=True), nn.Linear(DIM, DIM))
        self.norm = nn.LayerNorm(DIM, eps=1e-3)
        self.agg_kk = SoftAgg(DIM)
        self.agg_ij = SoftAgg(DIM)
        self.gru = nn.Sequential(nn.LayerNorm(DIM, eps=1e-3), GatedResidual(DIM),
                                 nn.LayerNorm(DIM, eps=1e-3), GatedResidual(DIM))
        self.corr = nn.Sequential(nn.Linear(2 * 49 * p * p, DIM), nn.ReLU(inplace=True),
                                  nn.Linear(DIM, DIM), nn.LayerNorm(DIM, eps=1e-3),
                                  nn.ReLU(inplace=True), nn.Linear(DIM, DIM))
        # reference: Sequential(ReLU, Linear, GradientClip[, Sigmoid]); GradientClip is identity in forward
        self.d = nn.Sequential(nn.ReLU(inplace=False), nn.Linear(DIM, 2), nn.Identity())
        self.w = nn.Sequential(nn.ReLU(inplace=False), nn.Linear(DIM, 2), nn.Identity(), nn.Sigmoid())

    @staticmethod
    def _soft_agg(agg, net, plan, E):
        """net += h( sum_g f(x) * softmax_g(g(x)) )[group]  (blocks.py:42-48, net.py:84-85)"""
        L = _lib.lib()
        x = net[0]
        fx = agg.f(x).contiguous()
        gx = agg.g(x).contiguous()
        y = torch.zeros(E, x.shape[1], dtype=fx.dtype, device=x.device)   # rows >= #groups stay 0
        st = _lib.stream_ptr(x.device)
        _lib.check(L.rvo_softagg(_lib.ptr(fx), _lib.ptr(gx), _lib.dtype_code(fx), _lib.ptr(plan), E,
                                 x.shape[1], 0, _lib.ptr(y), _lib.dtype_code(y), st), "rvo_softagg")
        hy = agg.h(y).contiguous()
        _lib.check(L.rvo_expand_add(_lib.ptr(hy), _lib.dtype_code(hy), _lib.ptr(plan), E, x.shape[1],
                                    _lib.ptr(x), st), "rvo_expand_add")

    # ---- mixed-precision fused path -------------------------------------------------------
    def _fused_weights(self):
        """fp16 copies of the Linear weights (concatenated where two layers share an input) and
        fp32 LayerNorm / head parameters, cached for inference."""
        # keyed on every parameter's storage and version: load_state_dict / .to() / an optimiser step rebuild it
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        w = getattr(self, "_fw", None)
        if w is not None and getattr(self, "_fw_key", None) == key:
            return w
        self._fw_key = key
        h = lambda t: t.detach().to(torch.float16).contiguous()
        f = lambda t: t.detach().float().contiguous()
        w = {}
        for name, lin in (("corr0", self.corr[0]), ("corr2", self.corr[2]), ("corr5", self.corr[5]),
                          ("c1a", self.c1[0]), ("c1b", self.c1[2]), ("c2a", self.c2[0]), ("c2b", self.c2[2]),
                          ("kk_h", self.agg_kk.h), ("ij_h", self.agg_ij.h),
                          ("g1_gate", self.gru[1].gate[0]), ("g1_a", self.gru[1].res[0]), ("g1_b", self.gru[1].res[2]),
                          ("g3_gate", self.gru[3].gate[0]), ("g3_a", self.gru[3].res[0]), ("g3_b", self.gru[3].res[2])):
            w[name] = (h(lin.weight), h(lin.bias))
        for name, agg in (("kk_fg", self.agg_kk), ("ij_fg", self.agg_ij)):
            w[name] = (h(torch.cat([agg.f.weight, agg.g.weight], 0)), h(torch.cat([agg.f.bias, agg.g.bias], 0)))
        for name, ln in (("ln_corr", self.corr[3]), ("ln_norm", self.norm), ("ln_g0", self.gru[0]), ("ln_g2", self.gru[2])):
            w[name] = (f(ln.weight), f(ln.bias))
        w["corr0p"] = F.pad(w["corr0"][0], (0, 896 - w["corr0"][0].shape[1])).contiguous()   # K 882 -> 896
        cols, ref = altcorr.tile_layout_index(2)
        w0t = torch.zeros(w["corr0"][0].shape[0], 18 * altcorr.TILE_GROUP, dtype=torch.float16,
                          device=w["corr0"][0].device)
        w0t[:, cols.to(w0t.device)] = w["corr0"][0][:, ref.to(w0t.device)]
        w["corr0t"] = w0t                                          # same layer, tile-layout input
        # heads run in fp16 under autocast: keep fp16-rounded values, stored fp32 for the kernel
        w["d"] = (f(h(self.d[1].weight)), f(h(self.d[1].bias)))
        w["w"] = (f(h(self.w[1].weight)), f(h(self.w[1].bias)))
        self._fw = w
        return w

    def invalidate_cache(self):
        self._fw = None
        self._fw_key = None

    def _fused_ready(self):
        """the fused mixed-precision path needs CUDA parameters (it has no other requirement)"""
        return self.norm.weight.is_cuda

    def _forward_chains(self, net, inp, corr, ii, jj, kk, plans, net_out=None):
        """EXPERIMENTAL alternative to _forward_fused (Update.use_chains = True; off by default because it is slower:
        1.05 ms vs 0.77 ms at E = 45 888, profiles/r02_up_chain_experiment.md).
        Same arithmetic and dtypes as the reference under autocast in 9 launches: the five row-local stretches of
        net.py:69-90 are one rvo_up_chain each (csrc/up_chain.cu: the 128-row activation tile stays in shared memory
        from Linear to Linear, LayerNorm / residual / gate / heads in the epilogues), separated by the two neighbour
        gathers (folded into the next stretch's operand load) and the two SoftAgg reductions (rvo_up_softagg_fg +
        the small per-group Linear h).  inp: [1,E,384] tensor or (imap_table [N,384] fp16, idx, mod)."""
        L = _lib.lib()
        W = self._fused_weights()
        E = ii.numel()
        dev = net.device
        st = _lib.stream_ptr(dev)
        P = _lib.ptr
        f16 = lambda *shape: torch.empty(*(shape or (E, DIM)), dtype=torch.float16, device=dev)
        f32 = lambda: torch.empty(chain_rows(E), DIM, dtype=torch.float32, device=dev)    # tile-blocked layout
        addr = _addr
        chain = lambda prologue, layers, **kw: run_chain(E, prologue, layers, st, **kw)

        lay = lambda key, epi, **extra: (addr(W[key][0]), addr(W[key][1]), epi, extra)
        def lay_half(key, half, epi, **extra):          # f (rows 0..383) or g (rows 384..767) of a fused [768, 384] weight
            w, b = W[key]
            return (addr(w, half * DIM * DIM * 2), addr(b, half * DIM * 2), epi, extra)
        rows = int(L.rvo_up_chain_scratch_rows())
        scr = _lib.Workspace.get(dev, rows * DIM * 6 + 512, "up_chain")
        scr32 = ctypes.c_void_p(scr.data_ptr())
        scr16 = ctypes.c_void_p(scr.data_ptr() + rows * DIM * 4)

        # ---- stretch 1: net = norm(net + inp + corr-MLP(corr))                                 net.py:74-75
        K0 = corr.shape[-1]
        if K0 == 18 * altcorr.TILE_GROUP and corr.dtype == torch.float16:
            c0, w0 = corr.reshape(E, K0), W["corr0t"]          # tile layout of altcorr.corr_tiles, permuted weight columns
        elif corr.dtype == torch.float16 and corr.stride(-1) == 1 and corr.stride(-2) == 896 and K0 == 882:
            c0, w0 = torch.as_strided(corr, (E, 896), (896, 1), corr.storage_offset()), W["corr0p"]   # zero-padded rows
        else:
            c0, w0 = F.pad(corr.reshape(E, -1).to(torch.float16), (0, 896 - K0)).contiguous(), W["corr0p"]
        if isinstance(inp, tuple):
            table, idx, mod = inp
            table = table.reshape(-1, DIM)
        else:
            table, idx, mod = inp.reshape(E, DIM).to(torch.float16).contiguous(), torch.arange(E, device=dev), 0
        net_in = net.reshape(E, DIM).float().contiguous()
        xa32, xa16, xb32, xb16 = f32(), f16(), f32(), f16()
        chain(_lib.PRO_ROWS,
              [(addr(w0), addr(W["corr0"][1]), _lib.EPI_RELU, {"K": w0.shape[1]}),
               lay("corr2", _lib.EPI_LN_RELU, gamma=W["ln_corr"][0], beta=W["ln_corr"][1]),
               lay("corr5", _lib.EPI_ADD3_LN, gamma=W["ln_norm"][0], beta=W["ln_norm"][1])],
              a16=c0, lda=c0.stride(0), net_in=net_in, imap16=table, imap_idx=idx, imap_mod=int(mod),
              out32=xa32, out16=xa16)
        # ---- stretch 2: net += c1(mask_ix * net[:, ix])                                        net.py:78-81
        chain(_lib.PRO_ROWS, [lay("c1a", _lib.EPI_RELU), lay("c1b", _lib.EPI_RES)],
              a16=xa16, lda=DIM, gather=plans.ix, res32=xa32, out32=xb32, out16=xb16)
        # ---- stretch 3: net += c2(mask_jx * net[:, jx]); f(net), g(net) of agg_kk              net.py:82,84
        fg = f16(E, 2 * DIM)
        chain(_lib.PRO_ROWS,
              [lay("c2a", _lib.EPI_RELU), lay("c2b", _lib.EPI_RES),
               lay_half("kk_fg", 0, _lib.EPI_STORE16, y16=fg, ldy=2 * DIM),
               lay_half("kk_fg", 1, _lib.EPI_STORE16, y16=addr(fg, DIM * 2), ldy=2 * DIM)],
              a16=xb16, lda=DIM, gather=plans.jx, res32=xb32, out32=xa32)
        x32 = xa32                                           # the hidden state after both neighbour MLPs
        def soft_agg(plan, cap, kh):
            y = f16(cap, DIM)
            _lib.check(L.rvo_up_softagg_fg(P(fg), P(plan), E, DIM, cap, P(y), st), "rvo_up_softagg_fg")
            w, b = W[kh]
            hy = f16(cap, DIM)
            _lib.check(L.rvo_up_linear(P(y), DIM, P(w), P(b), cap, DIM, DIM, 0, P(hy), DIM, st), "rvo_up_linear")
            return hy
        hy_k = soft_agg(plans.plan_k, plans.cap_k, "kk_h")
        grp_k, grp_ij = plans.edge_groups()
        # ---- stretch 4: f, g of agg_ij on net + agg_kk(net)                                    net.py:84-85
        chain(_lib.PRO_EXPAND,
              [lay_half("ij_fg", 0, _lib.EPI_STORE16, y16=fg, ldy=2 * DIM),
               lay_half("ij_fg", 1, _lib.EPI_STORE16, y16=addr(fg, DIM * 2), ldy=2 * DIM)],
              x32=x32, hy_a=hy_k, grp_a=grp_k)
        hy_ij = soft_agg(plans.plan_ij, plans.cap_ij, "ij_h")
        # ---- stretch 5: net = gru(net + agg_kk + agg_ij); heads                                net.py:86-90
        out = net_out if net_out is not None else torch.empty(1, E, DIM, dtype=torch.float32, device=dev)
        delta = torch.empty(1, E, 2, dtype=torch.float32, device=dev)
        weight = torch.empty(1, E, 2, dtype=torch.float32, device=dev)
        chain(_lib.PRO_EXPAND_LN,
              [lay("g1_gate", _lib.EPI_GATE), lay("g1_a", _lib.EPI_RELU),
               lay("g1_b", _lib.EPI_GATED_LN, gamma=W["ln_g2"][0], beta=W["ln_g2"][1]),
               lay("g3_gate", _lib.EPI_GATE), lay("g3_a", _lib.EPI_RELU), lay("g3_b", _lib.EPI_GATED_HEADS)],
              x32=x32, hy_a=hy_k, grp_a=grp_k, hy_b=hy_ij, grp_b=grp_ij,
              pro_gamma=W["ln_g0"][0], pro_beta=W["ln_g0"][1], out32=out,
              Wd=W["d"][0], bd=W["d"][1], Ww=W["w"][0], bw=W["w"][1], delta=delta, weight=weight,
              scratch32=scr32, scratch16=scr16)
        return out, (delta, weight, None)

    def _forward_fused(self, net, inp, corr, ii, jj, kk, plans, net_out=None):
        """Same arithmetic and dtypes as the reference under autocast, 17 fp16 GEMMs + 13 fused
        kernels instead of ~250 launches.  inp: [1,E,384] tensor or (imap_table [N,384] fp16, idx, mod)."""
        L = _lib.lib()
        W = self._fused_weights()
        E = ii.numel()
        dev = net.device
        st = _lib.stream_ptr(dev)
        P = _lib.ptr
        def linear(x, k, relu):
            # nn.Linear (+ ReLU) on the tcgen05 GEMM of librampvo (rvo_up_linear, csrc/up_gemm.cu): weights
            # resident in shared memory, activations streamed by TMA, bias / ReLU / fp16 pack in the epilogue
            w, b = W[k]
            y = torch.empty(x.shape[0], w.shape[0], dtype=torch.float16, device=dev)
            _lib.check(L.rvo_up_linear(P(x), x.stride(0), P(w), P(b), x.shape[0], w.shape[1], w.shape[0],
                                       int(relu), P(y), y.stride(0), st), "rvo_up_linear")
            return y
        lin = lambda x, k: linear(x, k, False)
        lin_relu = lambda x, k: linear(x, k, True)
        f16 = lambda: torch.empty(E, DIM, dtype=torch.float16, device=dev)
        f32 = lambda: torch.empty(E, DIM, dtype=torch.float32, device=dev)

        K0 = corr.shape[-1]
        if K0 == 18 * altcorr.TILE_GROUP and corr.dtype == torch.float16:
            # tile layout of altcorr.corr_tiles: first-layer weight columns permuted to match
            c = corr.reshape(E, K0)
            h = torch._addmm_activation(W["corr0"][1], c, W["corr0t"].t())
        elif corr.dtype == torch.float16 and corr.stride(-1) == 1 and corr.stride(-2) == 896 and K0 == 882:
            c = torch.as_strided(corr, (E, 896), (896, 1), corr.storage_offset())   # zero-padded rows
            h = torch._addmm_activation(W["corr0"][1], c, W["corr0p"].t())
        else:
            c = corr.reshape(E, -1).to(torch.float16).contiguous()
            h = torch._addmm_activation(W["corr0"][1], c, W["corr0"][0].t())
        h = lin(h, "corr2")
        h3 = f16()
        _lib.check(L.rvo_up_ln_relu(P(h), P(W["ln_corr"][0]), P(W["ln_corr"][1]), E, DIM, P(h3), st), "rvo_up_ln_relu")
        h4 = lin(h3, "corr5")
        if isinstance(inp, tuple):
            table, idx, mod = inp
            table = table.reshape(-1, DIM)
        else:
            table, idx, mod = inp.reshape(E, DIM).to(torch.float16).contiguous(), torch.arange(E, device=dev), 0
        net_in = net.reshape(E, DIM).float().contiguous()
        x, x16 = f32(), f16()
        _lib.check(L.rvo_up_add3_ln(P(net_in), P(table), P(idx), mod, P(h4), P(W["ln_norm"][0]),
                                    P(W["ln_norm"][1]), E, DIM, P(x), P(x16), st), "rvo_up_add3_ln")
        for ka, kb, nbr in (("c1a", "c1b", plans.ix), ("c2a", "c2b", plans.jx)):
            # mask * net[:, nbr] (net.py:78-82) is the operand load of the first Linear: rows of the fp16 copy of
            # the hidden state gathered by cp.async, a zero row where there is no neighbour
            w, b = W[ka]
            g = f16()
            _lib.check(L.rvo_up_linear_gather(P(x16), DIM, P(nbr), P(w), P(b), E, DIM, DIM, 1, P(g), DIM, st),
                       "rvo_up_linear_gather")
            t = lin(g, kb)
            _lib.check(L.rvo_up_add_cast(P(x), P(t), E, DIM, P(x16), st), "rvo_up_add_cast")
        n32, n16 = f32(), f16()
        for kfg, kh, plan, cap, with_ln in (("kk_fg", "kk_h", plans.plan_k, plans.cap_k, False),
                                            ("ij_fg", "ij_h", plans.plan_ij, plans.cap_ij, True)):
            fg = lin(x16, kfg)                                              # [E, 768] = [f | g]
            y = torch.empty(cap, DIM, dtype=torch.float16, device=dev)
            _lib.check(L.rvo_up_softagg_fg(P(fg), P(plan), E, DIM, cap, P(y), st), "rvo_up_softagg_fg")
            hy = lin(y, kh)
            if with_ln:
                _lib.check(L.rvo_up_expand_add_ln(P(hy), P(plan), E, DIM, P(x), None, P(W["ln_g0"][0]),
                                                  P(W["ln_g0"][1]), P(n32), P(n16), st), "rvo_up_expand_add_ln")
            else:
                _lib.check(L.rvo_up_expand_add_ln(P(hy), P(plan), E, DIM, P(x), P(x16), None, None, None,
                                                  None, st), "rvo_up_expand_add_ln")
        # gru: LN (done) -> GatedResidual -> LN -> GatedResidual -> heads
        a = lin(n16, "g1_gate")
        r = lin(lin_relu(n16, "g1_a"), "g1_b")
        m32, m16 = f32(), f16()
        _lib.check(L.rvo_up_gated_tail(P(n32), P(a), P(r), E, DIM, 0, P(W["ln_g2"][0]), P(W["ln_g2"][1]),
                                       P(m32), P(m16), None, None, None, None, None, None, st), "rvo_up_gated_tail")
        a = lin(m16, "g3_gate")
        r = lin(lin_relu(m16, "g3_a"), "g3_b")
        out = net_out if net_out is not None else torch.empty(1, E, DIM, dtype=torch.float32, device=dev)
        delta = torch.empty(1, E, 2, dtype=torch.float32, device=dev)
        weight = torch.empty(1, E, 2, dtype=torch.float32, device=dev)
        _lib.check(L.rvo_up_gated_tail(P(m32), P(a), P(r), E, DIM, 1, None, None, P(out), None,
                                       P(W["d"][0]), P(W["d"][1]), P(W["w"][0]), P(W["w"][1]), P(delta),
                                       P(weight), st), "rvo_up_gated_tail")
        return out, (delta, weight, None)

    @staticmethod
    def _soft_agg_autograd(agg, x, key):
        """SoftAgg.forward (blocks.py:42-48) with index ops that autograd differentiates"""
        _, g = torch.unique(key, return_inverse=True)
        n = int(g.max().item()) + 1
        gx, fx = agg.g(x)[0], agg.f(x)[0]
        idx = g[:, None].expand_as(gx)
        mx = torch.full((n, gx.shape[1]), -float("inf"), dtype=gx.dtype, device=gx.device)
        mx = mx.scatter_reduce(0, idx, gx.detach(), reduce="amax", include_self=True)
        ex = (gx - mx[g]).exp()
        w = ex / torch.zeros(n, gx.shape[1], dtype=gx.dtype, device=gx.device).index_add_(0, g, ex)[g]
        y = torch.zeros(n, gx.shape[1], dtype=gx.dtype, device=gx.device).index_add_(0, g, fx * w)
        return agg.h(y)[g][None]

    def _forward_autograd(self, net, inp, corr, ii, jj, kk):
        """the training path (gradients recorded): the reference's op sequence (net.py:69-90) in tensor ops; only
        the neighbour lookup, which carries no gradient, runs on the graph-plan kernels"""
        if isinstance(inp, tuple):
            table, idx, mod = inp
            inp = table.reshape(-1, DIM)[(idx % mod) if mod else idx][None]
        net = net + inp + self.corr(corr)
        net = self.norm(net)
        ix, jx = fastba.neighbors(kk, jj)
        mask_ix = (ix >= 0).float().reshape(1, -1, 1)
        mask_jx = (jx >= 0).float().reshape(1, -1, 1)
        net = net + self.c1(mask_ix * net[:, ix])
        net = net + self.c2(mask_jx * net[:, jx])
        net = net + self._soft_agg_autograd(self.agg_kk, net, kk)
        net = net + self._soft_agg_autograd(self.agg_ij, net, ii * 12345 + jj)
        net = self.gru(net)
        return net, (self.d(net), self.w(net), None)

    def forward(self, net, inp, corr, flow, ii, jj, kk, plans=None, net_out=None):
        """net [1,E,384], inp [1,E,384], corr [1,E,882], ii/jj/kk [E] ->
        (net [1,E,384] fp32, (delta [1,E,2], weight [1,E,2], None)).  Extensions: `plans`, a
        GraphPlans built once per graph (built here when omitted); `inp` may be the tuple
        (imap_table, index, modulo) so that the context gather is fused; `net_out`, a contiguous
        fp32 [1,E,384] buffer for the new hidden state (fused path only).  Under autocast the fused
        mixed-precision path runs; otherwise the generic path in the tensors' own dtype."""
        _lib.require_cuda(net, corr, ii, jj, kk)
        E = ii.numel()
        if torch.is_grad_enabled() and (net.requires_grad or corr.requires_grad or
                                        any(p.requires_grad for p in self.parameters())):
            return self._forward_autograd(net, inp, corr, ii, jj, kk)
        if plans is None:
            plans = GraphPlans(ii, jj, kk)
        dev = net.device
        if torch.is_autocast_enabled() and E > 0:
            with torch.cuda.device(dev), torch.autocast("cuda", enabled=False):
                fwd = self._forward_chains if getattr(self, "use_chains", False) else self._forward_fused
                return fwd(net, inp, corr, ii, jj, kk, plans, net_out)
        if isinstance(inp, tuple):
            table, idx, mod = inp
            inp = table.reshape(-1, DIM)[(idx % mod) if mod else idx][None]
        L = _lib.lib()
        st = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            net = net.float() + inp.float() + self.corr(corr).float()
            net = self.norm(net).float().contiguous()                     # [1,E,384] fp32
            x = net[0]
            g = torch.empty(E, DIM, dtype=torch.float32, device=dev)
            for mlp, idx in ((self.c1, plans.ix), (self.c2, plans.jx)):
                _lib.check(L.rvo_gather_rows(_lib.ptr(x), _lib.ptr(idx), E, DIM, _lib.ptr(g),
                                             _lib.dtype_code(g), st), "rvo_gather_rows")
                x.add_(mlp(g))
            self._soft_agg(self.agg_kk, net, plans.plan_k, E)
            self._soft_agg(self.agg_ij, net, plans.plan_ij, E)
            net = self.gru(net)
            return net, (self.d(net), self.w(net), None)


class Patchifier(nn.Module):
    """ramp/net.py:93-203 (MultiScale input mode; event-biased or random patch selection)."""

    def __init__(self, channels_dim, patch_size=3, input_mode="MultiScale"):
        super().__init__()
        self.input_mode = input_mode
        self.P = patch_size
        evs, img = channels_dim
        if input_mode == "SingleScale":       # net.py:101-111
            self.encoder = MergerLSTMsceneEncoder(evs_ch_dim=evs, img_ch_dim=img, output_lstm_dim=15,
                                                  output_dim_f=128, output_dim_i=DIM, norm_fn_fmap="instance",
                                                  norm_fn_imap="none", kernel_size_superstate=1)
        elif input_mode == "MultiScale":      # net.py:112-123
            self.encoder = MultiScaleMergerDoubleNet(evs_ch_dim=evs, img_ch_dim=img, lstm_dim=16,
                                                     output_dim_f=128, output_dim_i=DIM,
                                                     norm_fn_fmap="instance", norm_fn_imap="none",
                                                     norm_superstate=False)
        else:
            raise ValueError(f"Invalid input mode: {input_mode}")

    def forward(self, input_, patches_per_image=80, reinit_hidden=False, disps=None, event_bias=False,
                gradient_bias=False, gmap_out=None):
        events, images, mask = input_
        mask_l = torch.as_tensor(mask).reshape(-1).tolist()       # host-side, like evaluate.py:163
        # the patch selection reads only the events: with a side stream it becomes a parallel branch of the
        # captured graph (top-k alone is ~100 us of a single-CTA kernel)
        side = getattr(self, "branch_stream", None)
        coords = None
        if side is not None and event_bias and all(mask_l) and events.is_cuda:
            cur = torch.cuda.current_stream(events.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                coords = coords_from_topk_events(events, patches_per_image, non_max_supp_rad=11)
        # fmap / 4, imap / 4 (net.py:152-153) are folded into the encoders' last 1x1 convolutions
        enc_dtype = getattr(self, "encoder_autocast", None)      # training: run the encoder in bf16 (train.py)
        import contextlib
        ctx = (torch.autocast("cuda", dtype=enc_dtype) if (enc_dtype is not None and torch.is_grad_enabled())
               else contextlib.nullcontext())
        with ctx:
            if self.input_mode == "SingleScale":   # net.py:141-145: no mask, the presence tests are data-dependent
                fmap, imap, _ = self.encoder(events=events, images=images, reinit_hidden=reinit_hidden,
                                             out_scale=0.25)
                mask_l = [True] * fmap.shape[1]
            else:
                fmap, imap = self.encoder(events=events, images=images, mask=mask, reinit_hidden=reinit_hidden,
                                          out_scale=0.25)
        if fmap.dtype == torch.bfloat16:
            fmap, imap = fmap.float(), imap.float()
        if coords is not None:
            cur.wait_stream(side)
            if not torch.cuda.is_current_stream_capturing():
                coords.record_stream(cur)
        if not any(mask_l):
            return None, None, None, None, None, None
        if events.shape[1] == len(mask_l) and not all(mask_l):
            events = events[:, [t for t, keep in enumerate(mask_l) if keep]]
        b, n, c, h, w = fmap.shape
        dev = fmap.device
        if coords is not None:
            pass
        elif event_bias:
            coords = coords_from_topk_events(events, patches_per_image, non_max_supp_rad=11)
        else:
            x = torch.randint(1, w - 1, size=[n, patches_per_image], device=dev)
            y = torch.randint(1, h - 1, size=[n, patches_per_image], device=dev)
            coords = torch.stack([x, y], dim=-1).float()
        P = self.P
        if gmap_out is not None:
            gmap = altcorr.patchify(fmap[0], coords, P // 2, out=gmap_out)
        else:
            gmap = altcorr.patchify(fmap[0], coords, P // 2).view(b, -1, 128, P, P)
        imap = altcorr.patchify(imap[0], coords, 0).view(b, -1, DIM, 1, 1)
        if disps is None:
            disps = torch.ones(b, n, h, w, device=dev)
        ys, xs = torch.meshgrid(torch.arange(h, device=dev, dtype=torch.float32),
                                torch.arange(w, device=dev, dtype=torch.float32), indexing="ij")
        grid = torch.stack([xs.expand(n, h, w), ys.expand(n, h, w), disps[0]], dim=1)   # [n,3,h,w]
        patches = altcorr.patchify(grid, coords, P // 2).view(b, -1, 3, P, P)
        index = torch.arange(n, device=dev).view(n, 1).repeat(1, patches_per_image).reshape(-1)
        clr = altcorr.patchify(images[0].float(), 4 * (coords + 0.5), 0).view(b, -1, 3)
        return fmap, gmap, imap, patches, index, clr


class CorrBlock:
    """ramp/net.py:206-229: training-time correlation lookup (edge dropout in the backward pass only)"""

    def __init__(self, fmap, gmap, radius=3, dropout=0.2, levels=(1, 4)):
        self.dropout, self.radius, self.levels = dropout, radius, list(levels)
        self.gmap = gmap
        b, n, c, h, w = fmap.shape
        self.pyramid = [F.avg_pool2d(fmap.view(b * n, c, h, w), l, stride=l).view(b, n, c, h // l, w // l)
                        for l in self.levels]                                      # utils.py:81-90 pyramidify

    def __call__(self, ii, jj, coords):
        corrs = [altcorr.corr(self.gmap, self.pyramid[i], coords / self.levels[i], ii, jj, self.radius, self.dropout)
                 for i in range(len(self.levels))]
        return torch.stack(corrs, -1).view(1, len(ii), -1)


def motion_bootstrap(n, poses, MOTION_MODEL, MOTION_DAMPING):
    """ramp/pose_prediction/pose_pred_utils.py:189-198: damped-linear extrapolation of the newest pose"""
    if MOTION_MODEL == 'DAMPED_LINEAR':
        P1, P2 = SE3(poses[n - 1]), SE3(poses[n - 2])
        return (SE3.exp(MOTION_DAMPING * (P1 * P2.inv()).log()) * P1).data
    return poses[n - 1]


class VONet(nn.Module):
    """ramp/net.py:232-249 (constructor surface).  `.forward` of the reference is the training
    unroll (net.py:252-378), which belongs to the training row."""

    def __init__(self, cfg):
        super().__init__()
        self.P = 3
        self.RES = 4
        self.DIM = DIM
        self.EVENT_BIAS = cfg["event_bias"]
        self.MOTION_MODEL = "DAMPED_LINEAR"
        self.MOTION_DAMPING = 0.5
        self.inp_channel_dims = get_channel_dim(cfg)
        self.input_mode = cfg["input_mode"]
        self.patchify = Patchifier(channels_dim=self.inp_channel_dims, patch_size=self.P,
                                   input_mode=self.input_mode)
        self.update = Update(self.P)

    def forward(self, input_, poses, disps, intrinsics, STEPS=12, structure_only=False):
        """The training unroll (ramp/net.py:252-378): patchify a whole clip, then STEPS recurrent updates, each
        followed by two differentiable Gauss-Newton steps (rampvo_b200.ba.BA); frames beyond the first 8 join one
        per step.  Returns the reference's trajectory list [(valid, coords, coords_gt, Gs[:, :n], Ps[:, :n])].
        (Upstream unpacks five values from patchify, which returns six — net.py:263 vs :203 — so the reference
        cannot run this method as shipped; the unpack is the only deviation here.)"""
        input_ = preprocess_input(input_tensor=input_)
        if not isinstance(poses, SE3):
            poses = SE3(poses)
        intrinsics = intrinsics / 4.0
        disps = disps[:, :, 1::4, 1::4].float()
        fmap, gmap, imap, patches, ix, _ = self.patchify(input_=input_, disps=disps, reinit_hidden=True,
                                                         event_bias=self.EVENT_BIAS)
        corr_fn = CorrBlock(fmap, gmap)
        b, N, c, h, w = fmap.shape
        p = self.P
        dev = fmap.device
        patches_gt = patches.clone()
        Ps = poses
        d = patches[..., 2, p // 2, p // 2]
        patches = patches.clone()
        patches[..., 2, :, :] = torch.rand_like(d)[..., None, None]                  # set_depth, utils.py:99-101
        kk, jj = flatmeshgrid(torch.where(ix < 8)[0], torch.arange(0, 8, device=dev), indexing="ij")
        ii = ix[kk]
        imap = imap.view(b, -1, DIM)
        net = torch.zeros(b, len(kk), DIM, device=dev, dtype=torch.float)
        Gs = SE3.IdentityLike(poses)
        if structure_only:
            Gs = SE3(poses.data.clone())
        traj = []
        bounds = [-64, -64, w + 64, h + 64]
        n_input = input_[1].shape[1]
        while len(traj) < STEPS:
            Gs = Gs.detach()
            patches = patches.detach()
            n = int(ii.max().item()) + 1
            if len(traj) >= 8 and n < n_input:
                if not structure_only:
                    data = Gs.data.clone()
                    data[:, n] = motion_bootstrap(MOTION_DAMPING=self.MOTION_DAMPING, MOTION_MODEL=self.MOTION_MODEL,
                                                  poses=Gs.data[0, :], n=n)
                    Gs = SE3(data)
                kk1, jj1 = flatmeshgrid(torch.where(ix < n)[0], torch.arange(n, n + 1, device=dev), indexing="ij")
                kk2, jj2 = flatmeshgrid(torch.where(ix == n)[0], torch.arange(0, n + 1, device=dev), indexing="ij")
                ii = torch.cat([ix[kk1], ix[kk2], ii])
                jj = torch.cat([jj1, jj2, jj])
                kk = torch.cat([kk1, kk2, kk])
                net = torch.cat([torch.zeros(b, len(kk1) + len(kk2), DIM, device=dev), net], dim=1)
                if np.random.rand() < 0.1:
                    k = (ii != (n - 4)) & (jj != (n - 4))
                    ii, jj, kk, net = ii[k], jj[k], kk[k], net[:, k]
                patches = patches.clone()
                patches[:, ix == n, 2] = torch.median(patches[:, (ix == n - 1) | (ix == n - 2), 2])
                n = int(ii.max().item()) + 1
            coords = pops.transform(Gs, patches, intrinsics, ii, jj, kk)
            coords1 = coords.permute(0, 1, 4, 2, 3).contiguous()
            corr = corr_fn(kk, jj, coords1)
            net, (delta, weight, _) = self.update(net, imap[:, kk], corr, None, ii, jj, kk)
            target = coords[..., p // 2, p // 2, :] + delta
            for _ in range(2):
                Gs, patches = BA(Gs, patches, intrinsics, target, weight, 1e-4, ii, jj, kk, bounds, ep=10,
                                 fixedp=1, structure_only=structure_only)
            dij = (ii - jj).abs()
            k = (dij > 0) & (dij <= 2)
            coords = pops.transform(Gs, patches, intrinsics, ii[k], jj[k], kk[k])
            coords_gt, valid, _ = pops.transform(Ps, patches_gt, intrinsics, ii[k], jj[k], kk[k], jacobian=True)
            traj.append((valid, coords, coords_gt, Gs[:, :n], Ps[:, :n]))
        return traj
