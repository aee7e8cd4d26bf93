"""CPU restatement (PyTorch, fp32) of the reference's per-frame path — TEST / BASELINE INFRASTRUCTURE.

This is the "reference CPU-only path (PyTorch conv + Python BA, CUDA ops bypassed)" that
BASELINE.json asks to be timed beside the B200 numbers: the reference itself cannot run on a CPU
(device="cuda" is hard-coded, ramp/net.py:175-199, ramp/Ramp_vo.py:24-94, and its three extensions are
CUDA-only / need Eigen), so every stage is restated with the plumbing the reference would use:

  encoder     ramp/extractor.py:468-566 — strided conv + nn.LSTM over B*H*W per-pixel sequences
              (:351-381, including the permute/contiguous round trips), 1x1-conv super state, two
              residual CNNs with InstanceNorm / no norm (:8-57, :272-311)
  patchify    ramp/altcorr/correlation_kernel.cu:17-47 + correlation.py:51-68 as gather + blend
  reproject   ramp/projective_ops.py:16-101
  corr        ramp/altcorr/correlation_kernel.cu:83-136,221-232, ramp/Ramp_vo.py:175-182 as
              index + einsum over 128 channels per 8x8 window, chunked over edges
  update      ramp/net.py:69-90, ramp/blocks.py:15-50 with scatter_softmax / scatter_sum written
              with index ops (torch_scatter is absent), neighbors per ramp/fastba/ba.cpp:59-97
  BA          ramp/ba.py:86-182 (the reference's Python BA) with ep = 1 to match cuda_ba's damping
Only tests/ and bench.py (cpu_baseline / --impl reference) import this module.  The numeric oracle
for the CUDA kernels is oracle/ref_ops.py; this file adds the network stages and the timing path.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ref_ops as O

DIM = 384


# --------------------------------------------------------------------------- encoder

def _res_block(x, p, pre, instance, stride):
    """ramp/extractor.py:47-57"""
    norm = F.instance_norm if instance else (lambda t: t)
    y = F.relu(norm(F.conv2d(x, p[pre + "conv1.weight"], p[pre + "conv1.bias"], stride=stride, padding=1)))
    y = F.relu(norm(F.conv2d(y, p[pre + "conv2.weight"], p[pre + "conv2.bias"], padding=1)))
    if stride != 1:
        x = norm(F.conv2d(x, p[pre + "downsample.0.weight"], p[pre + "downsample.0.bias"], stride=stride))
    return F.relu(x + y)


def _cnn(x, x2, x4, p, pre, instance):
    """MultiScaleBasicEncoder4.forward, ramp/extractor.py:288-311"""
    x = F.conv2d(x, p[pre + "conv1.weight"], p[pre + "conv1.bias"], stride=2, padding=3)
    if instance:
        x = F.instance_norm(x)
    x = F.relu(x)
    x = _res_block(x, p, pre + "layer1.0.", instance, 1)
    x = _res_block(x, p, pre + "layer1.1.", instance, 1)
    x = torch.cat((x, x2), dim=1)
    x = _res_block(x, p, pre + "layer3.0.", instance, 2)
    x = _res_block(x, p, pre + "layer3.1.", instance, 1)
    x = torch.cat((x, x4), dim=1)
    return F.conv2d(x, p[pre + "conv3.weight"], p[pre + "conv3.bias"])


class Encoder:
    """MultiScaleMergerDoubleNet (ramp/extractor.py:468-566) driven from a state dict with the
    reference's key names (prefix 'patchify.encoder.')."""

    def __init__(self, state_dict, prefix="patchify.encoder."):
        self.p = {k[len(prefix):]: v.detach().float().cpu() for k, v in state_dict.items()
                  if k.startswith(prefix)}
        self.lstm = {}
        for kind, cin in (("ev_encoders", 5), ("im_encoders", 3)):
            for s in range(3):
                pre = "%s.%d.convlstm." % (kind, s)
                hid = self.p[pre + "weight_hh_l0"].shape[1]
                m = nn.LSTM(input_size=self.p[pre + "weight_ih_l0"].shape[1], hidden_size=hid, batch_first=True)
                m.load_state_dict({k[len(pre):]: v for k, v in self.p.items() if k.startswith(pre)})
                self.lstm[(kind, s)] = m.eval()
        self.super_states = [None, None, None]

    def _lstm_encoder(self, x, kind, s):
        """LSTMEncoder.forward, ramp/extractor.py:383-385,364-381"""
        scale = (1, 2, 4)[s]
        pre = "%s.%d.conv_1." % (kind, s)
        if scale <= 1:
            x = F.conv2d(x, self.p[pre + "weight"], self.p[pre + "bias"])
        else:
            x = F.conv2d(x, self.p[pre + "weight"], self.p[pre + "bias"], stride=scale, padding=1)
        T, C, H, W = x.shape
        seq = x[None].permute(0, 3, 4, 1, 2).contiguous().view(H * W, T, C)        # to_sequence
        out, _ = self.lstm[(kind, s)](seq)
        return out.view(1, H, W, T, -1).permute(0, 3, 4, 1, 2)[0]                   # [T,h,H,W]

    def __call__(self, events, images, mask, reinit_hidden=False):
        with torch.no_grad():
            ss_all = []
            for s in range(3):
                if reinit_hidden:
                    self.super_states[s] = None
                he = self._lstm_encoder(events[0], "ev_encoders", s)
                hi = self._lstm_encoder(images[0], "im_encoders", s)
                ss = self.super_states[s]
                outs, n_im = [], 0
                for t in range(he.shape[0]):                                           # forward_superstate
                    for kind, data, use in (("super_state_ev_encoder", he[t], True),
                                            ("super_state_im_encoders", hi[min(n_im, hi.shape[0] - 1)], bool(mask[t]))):
                        if not use:
                            continue
                        prev = torch.zeros_like(data) if ss is None else ss
                        pre = "%s.%d.encoder." % (kind, s)
                        ss = F.conv2d(torch.cat((prev, data), dim=0)[None], self.p[pre + "weight"],
                                      self.p[pre + "bias"])[0]
                    if mask[t]:
                        n_im += 1
                        outs.append(ss)
                allss = torch.stack(outs) if outs else ss[None]
                self.super_states[s] = allss[-1]
                ss_all.append(allss)
            fmap = _cnn(ss_all[0], ss_all[1], ss_all[2], self.p, "fmap_encoder.", True)
            imap = _cnn(ss_all[0], ss_all[1], ss_all[2], self.p, "imap_encoder.", False)
            return fmap[None], imap[None]


# --------------------------------------------------------------------------- update operator

def _lin(x, p, pre):
    return F.linear(x, p[pre + "weight"], p[pre + "bias"])


def _ln(x, p, pre):
    return F.layer_norm(x, (x.shape[-1],), p[pre + "weight"], p[pre + "bias"], eps=1e-3)


def _mlp2(x, p, pre):
    return _lin(F.relu(_lin(x, p, pre + "0.")), p, pre + "2.")


def _soft_agg(x, key, p, pre):
    """SoftAgg.forward, ramp/blocks.py:42-48"""
    _, jx = torch.unique(key, return_inverse=True)
    n = int(jx.max()) + 1
    gx, fx = _lin(x, p, pre + "g."), _lin(x, p, pre + "f.")
    idx = jx[:, None].expand_as(gx)
    mx = torch.full((n, gx.shape[1]), -float("inf")).scatter_reduce(0, idx, gx, reduce="amax")
    ex = (gx - mx[jx]).exp()
    w = ex / torch.zeros(n, gx.shape[1]).index_add_(0, jx, ex)[jx]
    y = torch.zeros(n, gx.shape[1]).index_add_(0, jx, fx * w)
    return _lin(y, p, pre + "h.")[jx]


def _gated(x, p, pre):
    return x + torch.sigmoid(_lin(x, p, pre + "gate.0.")) * _mlp2(x, p, pre + "res.")


def update_operator(p, net, inp, corr, ii, jj, kk):
    """Update.forward, ramp/net.py:69-90.  p: state dict with prefix 'update.' stripped; tensors [E,*]."""
    with torch.no_grad():
        c = _lin(F.relu(_lin(corr, p, "corr.0.")), p, "corr.2.")
        c = _lin(F.relu(_ln(c, p, "corr.3.")), p, "corr.5.")
        net = _ln(net + inp + c, p, "norm.")
        ix, jx = O.neighbors(kk.numpy(), jj.numpy())
        ix, jx = torch.from_numpy(ix), torch.from_numpy(jx)
        net = net + _mlp2((ix >= 0).float()[:, None] * net[ix], p, "c1.")
        net = net + _mlp2((jx >= 0).float()[:, None] * net[jx], p, "c2.")
        net = net + _soft_agg(net, kk, p, "agg_kk.")
        net = net + _soft_agg(net, ii * 12345 + jj, p, "agg_ij.")
        net = _gated(_ln(net, p, "gru.0."), p, "gru.1.")
        net = _gated(_ln(net, p, "gru.2."), p, "gru.3.")
        d = _lin(F.relu(net), p, "d.1.")
        w = torch.sigmoid(_lin(F.relu(net), p, "w.1."))
        return net, d, w


# --------------------------------------------------------------------------- corr (vectorised)

def corr_pyramid_torch(gmap, pyramid, coords, kk, jj, radius=3, chunk=2048):
    """Ramp_vo.corr (ramp/Ramp_vo.py:175-182) with torch ops: gmap [Np,C,P,P], pyramid list of
    [Nf,C,H,W], coords [E,2,P,P], kk/jj already reduced modulo the ring sizes -> [E, 882]."""
    E, P = coords.shape[0], coords.shape[-1]
    D = 2 * radius + 2
    d = D - 1
    offs = torch.arange(D) - radius
    outs = []
    for lvl, fm in enumerate(pyramid):
        Nf, C, H, W = fm.shape
        # pixel-major frame table: one window tap = one contiguous C-vector (a CPU gathers rows, not scalars)
        table = fm.permute(0, 2, 3, 1).reshape(Nf * H * W, C).contiguous()
        c = coords / (1, 4)[lvl]
        res = torch.empty(E, d, d, P, P)
        for s in range(0, E, chunk):
            cc = c[s:s + chunk]
            n = cc.shape[0]
            fx, fy = cc[:, 0].floor(), cc[:, 1].floor()
            dx, dy = (cc[:, 0] - fx)[:, None, None], (cc[:, 1] - fy)[:, None, None]
            xs = fx.long()[:, :, :, None, None] + offs.view(1, 1, 1, 1, D)    # [n,P,P,1,D] (columns)
            ys = fy.long()[:, :, :, None, None] + offs.view(1, 1, 1, D, 1)    # [n,P,P,D,1] (rows)
            ok = ((xs >= 0) & (xs < W) & (ys >= 0) & (ys < H))
            lin = (ys.clamp(0, H - 1) * W + xs.clamp(0, W - 1))
            lin = lin + (jj[s:s + chunk] * (H * W)).view(n, 1, 1, 1, 1)
            win = table.index_select(0, lin.reshape(-1)).view(n * P * P, D * D, C)
            g = gmap[kk[s:s + chunk]].permute(0, 2, 3, 1).reshape(n * P * P, C, 1)
            raw = (torch.bmm(win, g).view(n, P, P, D, D) * ok).permute(0, 3, 4, 1, 2)   # [n,a,b,i,j]
            o = ((1 - dx) * (1 - dy) * raw[:, :d, :d] + dx * (1 - dy) * raw[:, :d, 1:] +
                 (1 - dx) * dy * raw[:, 1:, :d] + dx * dy * raw[:, 1:, 1:])
            res[s:s + n] = o.permute(0, 2, 1, 3, 4)
        outs.append(res)
    return torch.stack(outs, -1).reshape(E, -1)


# --------------------------------------------------------------------------- one frame on the CPU

def python_ba(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations=2):
    """ramp/ba.py:86-182 restated on numpy arrays (same Schur algebra as cuda_ba with ep = 1)."""
    return O.ba(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations,
                dtype=np.float32)


def cpu_update_step(p_update, prob, gmap, pyramid, imap, net, n_edges=None):
    """reproject -> corr -> update operator -> 2 BA iterations on the first `n_edges` edges of a
    synthetic steady-state problem (rampvo_b200.synth.make_problem).  Returns the wall time split."""
    import time
    E = prob["E"] if n_edges is None else min(n_edges, prob["E"])
    ii, jj, kk = prob["ii"][:E], prob["jj"][:E], prob["kk"][:E]
    tm = {}
    t = time.perf_counter()
    coords = O.transform(prob["poses"], prob["patches"], prob["intrinsics"], ii, jj, kk, dtype=np.float32)[0]
    coords_t = torch.from_numpy(np.ascontiguousarray(coords.transpose(0, 3, 1, 2))).float()
    tm["reproject"] = time.perf_counter() - t
    t = time.perf_counter()
    M = prob["M"]
    corr = corr_pyramid_torch(gmap, pyramid, coords_t, torch.from_numpy(kk % (M * 32)),
                              torch.from_numpy(jj % 32))
    tm["corr"] = time.perf_counter() - t
    t = time.perf_counter()
    it, jt, kt = torch.from_numpy(ii), torch.from_numpy(jj), torch.from_numpy(kk)
    net2, d, w = update_operator(p_update, net[:E], imap[kt % (M * 32)], corr, it, jt, kt)
    tm["update_op"] = time.perf_counter() - t
    t = time.perf_counter()
    target = coords[:, 1, 1] + d.numpy()
    python_ba(prob["poses"], prob["patches"], prob["intrinsics"], target, w.numpy(), 1e-4, ii, jj, kk,
              prob["t0"], prob["t1"], 2)
    tm["ba"] = time.perf_counter() - t
    return tm
