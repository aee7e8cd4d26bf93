"""RAMP encoders (ramp/extractor.py) — same parameter tree as the reference so its checkpoints load
unchanged (state-dict keys are the compatibility surface), forward pass
restructured for one (event voxel, image) pair per call:

  * the per-pixel nn.LSTM of LSTMEncoder (extractor.py:351-381) sees sequences of length 1 with a
    zero initial state (MultiScale never passes `hx`, :378), i.e. it is the gated map
        h = sigmoid(o) * tanh(sigmoid(i) * tanh(g)),  [i,f,g,o] = W_ih x + b_ih + b_hh,
    applied per pixel; no permute/contiguous round trips, no cuDNN RNN launch with batch = H*W;
  * everything runs channels-last (NHWC) so the fmap lands directly in the layout the altcorr
    tensor-core path reads, and the level-2 pyramid entry is produced with it;
  * the dense 7x7 / 3x3 / 1x1 convolutions run on the hand-written tcgen05 implicit-GEMM kernel
    (rvo_conv2d_nhwc, csrc/conv_tc.cu): resident weights, im2col gather by cp.async, the channel concatenations
    read from their two sources, InstanceNorm statistics in the epilogue — no cuDNN on the fast path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

DIM = 32
CL = torch.channels_last


def _fast(x):
    """The fused channels-last fp16 path: CUDA, inference, mixed precision (Ramp_vo.py:23,331)."""
    return x.is_cuda and torch.is_autocast_enabled() and not torch.is_grad_enabled()


class _StatArena:
    """InstanceNorm statistics of every convolution of a forward pass live in one zeroed buffer: the kernels
    accumulate into their slice, so a frame needs ONE memset (begin()) instead of one per layer.  Outside a
    begin()/end() bracket each request gets its own zeroed tensor."""
    _cur = None

    @classmethod
    def begin(cls, device, floats=8192):
        cls._cur = [torch.zeros(floats, dtype=torch.float32, device=device), 0]

    @classmethod
    def end(cls):
        cls._cur = None

    @classmethod
    def take(cls, device, n):
        c = cls._cur
        if c is None or c[0].device != device or c[1] + n > c[0].numel():
            return torch.zeros(n, dtype=torch.float32, device=device)
        out = c[0][c[1]:c[1] + n]
        c[1] += n
        return out


def _packed_conv(conv, scale, pad_in):
    """fp16 [Cout, Kpad] weight matrix in the K order of rvo_conv2d_nhwc (tap-major, then input channel; the input
    channels optionally zero-padded to `pad_in`) + fp32 bias, both times `scale`; cached on the module and rebuilt
    whenever the parameters are replaced or modified in place (load_state_dict, .to(), optimiser steps)."""
    w, b = conv.weight, conv.bias
    key = (w.data_ptr(), w._version, b.data_ptr() if b is not None else 0, b._version if b is not None else 0,
           scale, pad_in)
    c = getattr(conv, "_wtc", None)
    if c is not None and c[0] == key:
        return c[1], c[2], c[3]
    Cout, Cin, ks, _ = w.shape
    wt = w.detach().float() * scale
    if pad_in and pad_in > Cin:
        wt = F.pad(wt, (0, 0, 0, 0, 0, pad_in - Cin))
        Cin = pad_in
    kpad = _lib.lib().rvo_conv2d_kpad(ks, Cin)
    if kpad < 0:
        raise RuntimeError("rvo_conv2d_nhwc: %d input channels (must be a multiple of 8)" % Cin)
    wp = torch.zeros(Cout, kpad, dtype=torch.float16, device=w.device)
    wp[:, :ks * ks * Cin] = wt.permute(0, 2, 3, 1).reshape(Cout, -1).half()
    bias = (b.detach().float() * scale).contiguous() if b is not None else None
    conv._wtc = (key, wp, bias, Cin)
    return wp, bias, Cin


def _conv_tc(conv, x, x2=None, scale=1.0, stats=False, pad_in=0):
    """nn.Conv2d on the tcgen05 implicit-GEMM kernel (rvo_conv2d_nhwc, csrc/conv_tc.cu).  x [1,C0,H,W] (and x2
    [1,C1,H,W]: the convolution sees their channel concatenation without materialising it) fp16 channels-last ->
    ([1,Cout,Ho,Wo] fp16 channels-last, InstanceNorm statistics [2*Cout] or None).  `scale` (a power of two) is
    folded into the packed weights and bias."""
    wp, bias, Cin = _packed_conv(conv, scale, pad_in)
    for t in (x, x2):
        if t is not None and not (t.dtype == torch.float16 and t.shape[0] == 1 and t.is_contiguous(memory_format=CL)):
            raise RuntimeError("rvo_conv2d_nhwc: inputs must be [1,C,H,W] fp16 channels-last")
    C0, C1 = x.shape[1], (x2.shape[1] if x2 is not None else 0)
    if C0 + C1 != Cin:
        raise RuntimeError("rvo_conv2d_nhwc: %d + %d input channels, the layer has %d" % (C0, C1, Cin))
    H, W = x.shape[-2:]
    ks, sd, pd = conv.kernel_size[0], conv.stride[0], conv.padding[0]
    Cout = conv.out_channels
    Ho, Wo = (H + 2 * pd - ks) // sd + 1, (W + 2 * pd - ks) // sd + 1
    out = torch.empty(1, Cout, Ho, Wo, dtype=torch.float16, device=x.device, memory_format=CL)
    st = _StatArena.take(x.device, 2 * Cout) if stats else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().rvo_conv2d_nhwc(_lib.ptr(x), C0, _lib.ptr(x2), C1, H, W, ks, sd, pd, _lib.ptr(wp),
                                              _lib.ptr(bias), Cout, _lib.ptr(out), _lib.ptr(st),
                                              _lib.stream_ptr(x.device)), "rvo_conv2d_nhwc")
    return out, st


def _stats(t):
    """InstanceNorm2d statistics of a [1,C,H,W] channels-last fp16 tensor -> device [2C] sums."""
    C = t.shape[1]
    sums = torch.empty(2 * C, dtype=torch.float32, device=t.device)
    _lib.check(_lib.lib().rvo_in_stats(_lib.ptr(t), t.shape[2] * t.shape[3], C, _lib.ptr(sums),
                                       _lib.stream_ptr(t.device)), "rvo_in_stats")
    return sums


def _apply(t, st=None, res=None, sr=None):
    """relu( [IN](res) + relu( [IN](t) ) ) in one pass (rvo_in_apply)."""
    out = torch.empty_like(t, memory_format=CL)
    _lib.check(_lib.lib().rvo_in_apply(_lib.ptr(t), _lib.ptr(st), _lib.ptr(res), _lib.ptr(sr),
                                       t.shape[2] * t.shape[3], t.shape[1], 1e-5, _lib.ptr(out),
                                       _lib.stream_ptr(t.device)), "rvo_in_apply")
    return out


class ResidualBlock(nn.Module):
    """extractor.py:8-57 (only the norms the hot path selects: 'instance' and 'none')."""

    def __init__(self, in_planes, planes, norm_fn='instance', stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, padding=1)
        self.instance = norm_fn == 'instance'
        if norm_fn not in ('instance', 'none'):
            raise NotImplementedError("norm_fn %r: the RAMP encoders use 'instance' / 'none'" % norm_fn)
        self.downsample = None
        if stride != 1:
            # reference: Sequential(Conv2d 1x1 stride, norm3); InstanceNorm2d has no parameters
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, kernel_size=1, stride=stride))

    def _norm(self, x):
        return F.instance_norm(x) if self.instance else x

    def _forward_fast(self, x, x2=None):
        """x (+ x2: channel concat) [1,C,H,W] fp16 channels-last.  Tensor-core conv with the InstanceNorm statistics
        in its epilogue -> normalise + ReLU (+ shortcut) in one pass, twice."""
        t, st = _conv_tc(self.conv1, x, x2, stats=self.instance)
        y = _apply(t, st)
        t, st = _conv_tc(self.conv2, y, stats=self.instance)
        if self.downsample is not None:
            d, sd = _conv_tc(self.downsample[0], x, x2, stats=self.instance)
            return _apply(t, st, d, sd)
        if x2 is not None:
            raise RuntimeError("ResidualBlock: a concatenated input needs the strided (downsample) variant")
        return _apply(t, st, x)

    def forward(self, x):
        if _fast(x) and x.shape[0] == 1:
            return self._forward_fast(x)
        y = F.relu(self._norm(self.conv1(x)))
        y = F.relu(self._norm(self.conv2(y)))
        if self.downsample is not None:
            x = self._norm(self.downsample(x))
        return F.relu(x + y)


class MultiScaleBasicEncoder4(nn.Module):
    """extractor.py:272-311 on top of BasicEncoder4 (:60-130).  `layer2` / `conv2` are created by
    the reference's double __init__ (:276-277) and never used; they are kept so that checkpoints
    load with strict=True."""

    def __init__(self, output_dim=128, norm_fn='instance', channel_dim=16, internal_input_dimensions=None):
        super().__init__()
        self.instance = norm_fn == 'instance'
        dims = internal_input_dimensions or [channel_dim] * 3
        self.conv1 = nn.Conv2d(channel_dim, DIM, kernel_size=7, stride=2, padding=3)
        self.layer1 = nn.Sequential(ResidualBlock(DIM, DIM, norm_fn, 1), ResidualBlock(DIM, DIM, norm_fn, 1))
        self.layer2 = nn.Sequential(ResidualBlock(DIM, 2 * DIM, norm_fn, 2),
                                    ResidualBlock(2 * DIM, 2 * DIM, norm_fn, 1))   # dead weights
        self.conv2 = nn.Conv2d(2 * DIM, 128, kernel_size=1)                          # dead weights
        c3 = DIM + dims[1]
        self.layer3 = nn.Sequential(ResidualBlock(c3, 2 * DIM, norm_fn, 2),
                                    ResidualBlock(2 * DIM, 2 * DIM, norm_fn, 1))
        self.conv3 = nn.Conv2d(2 * DIM + dims[2], output_dim, kernel_size=1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

    def forward(self, x, x_down2, x_down4, out_scale=1.0):
        """x [1,16,H,W], x_down2 [1,32,H/2,W/2], x_down4 [1,64,H/4,W/4] -> [1,out,H/4,W/4] (times
        out_scale: the /4 of net.py:152-153 folded into the last 1x1 conv on the fast path)."""
        if _fast(x) and x.shape[0] == 1:
            h16 = lambda t: t.half().contiguous(memory_format=CL)
            t, st = _conv_tc(self.conv1, h16(x), stats=self.instance)
            x = _apply(t, st)
            x = self.layer1(x)
            # torch.cat((x, x_down2)) / torch.cat((x, x_down4)) (extractor.py:302,309) are never materialised: the
            # convolutions read their K range from the two tensors
            x = self.layer3[0]._forward_fast(x, h16(x_down2))
            x = self.layer3[1](x)
            return _conv_tc(self.conv3, x, h16(x_down4), scale=out_scale)[0]
        x = self.conv1(x)
        if self.instance:
            x = F.instance_norm(x)
        x = F.relu(x)
        x = self.layer1(x)
        x = self.layer3(torch.cat((x, x_down2), dim=1))
        return self.conv3(torch.cat((x, x_down4), dim=1)) * out_scale


class BasicEncoder4(nn.Module):
    """extractor.py:60-130 (norm_fn 'instance' / 'none'): conv 7x7 s2 -> norm -> relu -> 2 ResBlocks(32) ->
    ResBlock(32->64, s2) + ResBlock(64) -> 1x1 conv.  InstanceNorm2d has no parameters, so the state-dict keys
    are the convolutions only, exactly like the reference's."""

    def __init__(self, output_dim=128, norm_fn='instance', channel_dim=15):
        super().__init__()
        if norm_fn not in ('instance', 'none'):
            raise NotImplementedError("norm_fn %r: the RAMP encoders use 'instance' / 'none'" % norm_fn)
        self.instance = norm_fn == 'instance'
        self.conv1 = nn.Conv2d(channel_dim, DIM, kernel_size=7, stride=2, padding=3)
        self.layer1 = nn.Sequential(ResidualBlock(DIM, DIM, norm_fn, 1), ResidualBlock(DIM, DIM, norm_fn, 1))
        self.layer2 = nn.Sequential(ResidualBlock(DIM, 2 * DIM, norm_fn, 2), ResidualBlock(2 * DIM, 2 * DIM, norm_fn, 1))
        self.conv2 = nn.Conv2d(2 * DIM, output_dim, kernel_size=1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

    def forward_fast(self, x16, out_scale=1.0):
        """x16 [1,16,H,W] fp16 channels-last (15 channels + a zero pad, see MergerLSTMsceneEncoder)"""
        t, st = _conv_tc(self.conv1, x16, stats=self.instance, pad_in=x16.shape[1])
        x = _apply(t, st)
        x = self.layer2(self.layer1(x))
        return _conv_tc(self.conv2, x, scale=out_scale)[0]

    def forward(self, x):
        """x [b,n,C,H,W] -> [b,n,out,H/4,W/4] (extractor.py:107-126)"""
        b, n, c, h, w = x.shape
        x = self.conv1(x.view(b * n, c, h, w))
        if self.instance:
            x = F.instance_norm(x)
        x = self.conv2(self.layer2(self.layer1(F.relu(x))))
        return x.view(b, n, *x.shape[1:])


class MergerLSTMsceneEncoder(nn.Module):
    """extractor.py:187-269, the SingleScale RAMP encoder (BASELINE.json configs[0]): two full-resolution per-pixel
    LSTMs whose (h, c) state is CARRIED across calls (:242-243), one shared 1x1 super-state convolution applied for
    the events and then for the image when they are non-zero (:253-258), two BasicEncoder4 CNNs."""

    def __init__(self, evs_ch_dim=5, img_ch_dim=3, output_lstm_dim=15, output_dim_f=128, output_dim_i=384,
                 norm_fn_fmap="instance", norm_fn_imap="none", kernel_size_superstate=1):
        super().__init__()
        if kernel_size_superstate != 1:
            raise NotImplementedError("kernel_size_superstate %d: every caller uses 1 (net.py:110)" % kernel_size_superstate)
        self.hidden_size = output_lstm_dim
        self.events_convlstm = nn.LSTM(input_size=evs_ch_dim, hidden_size=output_lstm_dim, batch_first=True)
        self.image_convlstm = nn.LSTM(input_size=img_ch_dim, hidden_size=output_lstm_dim, batch_first=True)
        self.superstate_encoder = nn.Conv2d(2 * output_lstm_dim, output_lstm_dim, kernel_size=1, padding=0)
        self.fmap_encoder = BasicEncoder4(output_dim_f, norm_fn_fmap, output_lstm_dim)
        self.imap_encoder = BasicEncoder4(output_dim_i, norm_fn_imap, output_lstm_dim)
        self.states_events, self.states_image, self.super_state = None, None, None
        self._dev_state = None          # fused path: (state_ev, state_im, super, flags) device buffers

    def reset_state(self):
        self.states_events, self.states_image, self.super_state = None, None, None
        self._dev_state = None

    def _packed_params(self, device):
        c = getattr(self, "_pp", None)
        if c is not None and c.device == device:
            return c
        parts = []
        for lstm in (self.events_convlstm, self.image_convlstm):
            parts += [lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0 + lstm.bias_hh_l0]
        h = self.hidden_size
        parts += [self.superstate_encoder.weight.reshape(h, 2 * h), self.superstate_encoder.bias]
        buf = torch.cat([t.detach().float().reshape(-1) for t in parts]).contiguous().to(device)
        n = _lib.lib().rvo_scene_lstm_params_floats(self.events_convlstm.input_size, self.image_convlstm.input_size)
        if n != buf.numel():
            raise RuntimeError("scene LSTM parameter block: %d floats, library expects %d" % (buf.numel(), n))
        self._pp = buf
        return buf

    def _forward_fast(self, events, images, reinit_hidden, out_scale):
        """one event stack + one image: rvo_scene_lstm_forward (2 launches) + the two channels-last CNNs"""
        ev = events[0, 0].float().contiguous()
        im = images[0, 0].float().contiguous()
        H, W = ev.shape[-2:]
        dev, h = ev.device, self.hidden_size
        st = self._dev_state
        first = reinit_hidden or st is None or st[0].shape[-1] != H * W
        if st is None or st[0].shape[-1] != H * W:
            st = self._dev_state = (torch.zeros(2, h, H * W, device=dev), torch.zeros(2, h, H * W, device=dev),
                                    torch.zeros(h, H * W, device=dev), torch.zeros(2, dtype=torch.int32, device=dev))
        out = torch.empty(1, 16, H, W, dtype=torch.float16, device=dev).contiguous(memory_format=CL)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().rvo_scene_lstm_forward(
                _lib.ptr(self._packed_params(dev)), ev.shape[0], im.shape[0], _lib.ptr(ev), _lib.ptr(im), H, W,
                _lib.ptr(st[0]), _lib.ptr(st[1]), _lib.ptr(st[2]), _lib.ptr(st[3]), int(first), _lib.ptr(out),
                _lib.stream_ptr(dev)), "rvo_scene_lstm_forward")
        _StatArena.begin(dev)
        fmap = self.fmap_encoder.forward_fast(out, out_scale)
        imap = self.imap_encoder.forward_fast(out, out_scale)
        _StatArena.end()
        return fmap[None], imap[None], None

    def forward(self, events, images, reinit_hidden=False, out_scale=1.0):
        """events [1,T,Ce,H,W], images [1,T,3,H,W] -> fmap [1,T,128,H/4,W/4], imap [1,T,384,H/4,W/4], lstm states"""
        if _fast(events) and events.shape[1] == 1 and images.shape[1] == 1 and events.shape[0] == 1:
            return self._forward_fast(events, images, reinit_hidden, out_scale)
        if reinit_hidden:
            self.states_events, self.states_image, self.super_state = None, None, None
        B, T, Ce, H, W = events.shape
        Ti, Ci = images.shape[1], images.shape[2]
        ev_seq = events.permute(0, 3, 4, 1, 2).contiguous().view(B * H * W, T, Ce)
        im_seq = images.permute(0, 3, 4, 1, 2).contiguous().view(B * H * W, Ti, Ci)
        oe, self.states_events = self.events_convlstm(ev_seq, self.states_events)
        oi, self.states_image = self.image_convlstm(im_seq, self.states_image)
        oe = oe.view(B, H, W, T, self.hidden_size).permute(0, 3, 4, 1, 2)
        oi = oi.view(B, H, W, Ti, self.hidden_size).permute(0, 3, 4, 1, 2)
        outs = []
        for t in range(min(T, Ti)):
            for data, present in ((oe[0, t], bool(torch.any(events[:, t] != 0))),
                                  (oi[0, t], bool(torch.any(images[:, t] != 0)))):
                if present:
                    prev = torch.zeros_like(data) if self.super_state is None else self.super_state
                    self.super_state = self.superstate_encoder(torch.cat((prev, data), dim=0)[None])[0]
            outs.append(self.super_state)
        ss = torch.stack(outs, dim=0)[None]
        return self.fmap_encoder(ss) * out_scale, self.imap_encoder(ss) * out_scale, [(oe, oi)]


class LSTMEncoder(nn.Module):
    """extractor.py:314-390: strided conv (k = s+1, stride s, pad 1; 1x1 at s <= 1) followed by the
    per-pixel LSTM cell evaluated for one step from a zero state."""

    def __init__(self, in_channels, downsample_scale=0, out_channels=15):
        super().__init__()
        k, s, p = downsample_scale + 1, downsample_scale, 1
        if downsample_scale <= 1:
            k, s, p = 1, 1, 0
        self.hidden = out_channels
        self.conv_1 = nn.Conv2d(in_channels, in_channels, kernel_size=k, stride=s, padding=p)
        self.convlstm = nn.LSTM(input_size=in_channels, hidden_size=out_channels, batch_first=True)

    def forward(self, x):
        """x [T,C,H,W] -> h [T,hidden,H/s,W/s].  The T frames of ONE call form a per-pixel sequence that starts from
        a zero state (extractor.py:364-381: to_sequence -> nn.LSTM -> from_sequence); online tracking feeds T = 1,
        where the recurrence collapses to a gated map of the input, training clips feed T = 15."""
        x = self.conv_1(x)
        T, C, H, W = x.shape
        if T > 1:
            seq = x.permute(2, 3, 0, 1).reshape(H * W, T, C)
            out, _ = self.convlstm(seq.to(self.convlstm.weight_ih_l0.dtype) if not torch.is_autocast_enabled() else seq)
            return out.reshape(H, W, T, self.hidden).permute(2, 3, 0, 1).to(x.dtype)
        w = self.convlstm.weight_ih_l0                       # [4h, C], gate order i, f, g, o
        b = self.convlstm.bias_ih_l0 + self.convlstm.bias_hh_l0
        gates = F.conv2d(x, w[:, :, None, None].to(x.dtype), b.to(x.dtype))
        i, _, g, o = gates.float().chunk(4, dim=1)
        c = torch.sigmoid(i) * torch.tanh(g)
        return (torch.sigmoid(o) * torch.tanh(c)).to(x.dtype)


class SuperStateEncoder(nn.Module):
    """extractor.py:393-412: 1x1 conv over cat(previous super state, new embedding)."""

    def __init__(self, kernel_size, out_channels=15, norm_superstate=False):
        super().__init__()
        self.encoder = nn.Conv2d(2 * out_channels, out_channels, kernel_size=kernel_size,
                                 padding=(kernel_size - 1) // 2)
        self.norm_superstate = norm_superstate

    def forward(self, data, prev_super_state=None):
        if prev_super_state is None:
            prev_super_state = torch.zeros_like(data)
        return self.encoder(torch.cat((prev_super_state, data), dim=1))


class MultiScaleMergerDoubleNet(nn.Module):
    """extractor.py:468-566.  Keeps one recurrent super state per scale across calls (the only
    cross-frame recurrence of the MultiScale encoder)."""

    def __init__(self, evs_ch_dim, img_ch_dim, lstm_dim=16, output_dim_f=128, output_dim_i=384,
                 norm_fn_fmap="instance", norm_fn_imap="none", kernel_size_superstate=1,
                 norm_superstate=False):
        super().__init__()
        self.scales = [1, 2, 4]
        self.ev_encoders = nn.ModuleList()
        self.im_encoders = nn.ModuleList()
        self.super_state_ev_encoder = nn.ModuleList()
        self.super_state_im_encoders = nn.ModuleList()
        dims = []
        for s in self.scales:
            h = lstm_dim * s
            dims.append(h)
            self.ev_encoders.append(LSTMEncoder(evs_ch_dim, s, h))
            self.im_encoders.append(LSTMEncoder(img_ch_dim, s, h))
            self.super_state_ev_encoder.append(SuperStateEncoder(kernel_size_superstate, h, norm_superstate))
            self.super_state_im_encoders.append(SuperStateEncoder(kernel_size_superstate, h, norm_superstate))
        self.super_states = [None, None, None]
        self.norm_superstate = norm_superstate
        self.fmap_encoder = MultiScaleBasicEncoder4(output_dim_f, norm_fn_fmap, lstm_dim, dims)
        self.imap_encoder = MultiScaleBasicEncoder4(output_dim_i, norm_fn_imap, lstm_dim, dims)

    def reset_state(self):
        self.super_states = [None, None, None]

    def _stem_params(self, k, device):
        """Packed fp32 parameters of scale k in the layout rvo_stem_params_layout reports."""
        cache = getattr(self, "_stem_cache", None)
        if cache is None:
            cache = self._stem_cache = {}
        if k in cache and cache[k][0].device == device:
            return cache[k]
        import ctypes
        ev, im = self.ev_encoders[k], self.im_encoders[k]
        se, si = self.super_state_ev_encoder[k].encoder, self.super_state_im_encoders[k].encoder
        h, Ce, Ci = ev.hidden, ev.conv_1.in_channels, im.conv_1.in_channels
        ks, st, pd = ev.conv_1.kernel_size[0], ev.conv_1.stride[0], ev.conv_1.padding[0]
        offs = (ctypes.c_int * 12)()
        total = ctypes.c_int(0)
        _lib.check(_lib.lib().rvo_stem_params_layout(Ce, Ci, ks, h, offs, ctypes.byref(total)),
                   "rvo_stem_params_layout")
        buf = torch.zeros(total.value, dtype=torch.float32)
        rows = lambda w: torch.cat([w[0:h], w[2 * h:3 * h], w[3 * h:4 * h]], 0)   # gates i, g, o
        parts = [ev.conv_1.weight, ev.conv_1.bias, im.conv_1.weight, im.conv_1.bias,
                 rows(ev.convlstm.weight_ih_l0), rows(ev.convlstm.bias_ih_l0 + ev.convlstm.bias_hh_l0),
                 rows(im.convlstm.weight_ih_l0), rows(im.convlstm.bias_ih_l0 + im.convlstm.bias_hh_l0),
                 se.weight.reshape(h, 2 * h).t(), se.bias, si.weight.reshape(h, 2 * h).t(), si.bias]
        for o, t in zip(offs, parts):
            t = t.detach().float().cpu().contiguous().reshape(-1)
            buf[o:o + t.numel()] = t
        cache[k] = (buf.to(device), (Ce, Ci, ks, st, pd, h))
        return cache[k]

    def _forward_fast(self, events, images, use_image, reinit_hidden, out_scale=1.0):
        """One event stack + (optionally) one image: the fused stem kernel per scale, then the two
        channels-last CNNs."""
        ev = events[0, 0].float().contiguous()
        im = images[0, 0].float().contiguous()
        H, W = ev.shape[-2:]
        L = _lib.lib()
        st = _lib.stream_ptr(ev.device)
        per_scale = []
        for k in range(3):
            if reinit_hidden:
                self.super_states[k] = None
            buf, (Ce, Ci, ks, sd, pd, h) = self._stem_params(k, ev.device)
            Ho, Wo = (H + 2 * pd - ks) // sd + 1, (W + 2 * pd - ks) // sd + 1
            prev = self.super_states[k]
            if prev is not None and not (prev.dtype == torch.float16 and prev.is_contiguous(memory_format=CL)):
                prev = prev.half().contiguous(memory_format=CL)
            out = torch.empty(1, h, Ho, Wo, dtype=torch.float16, device=ev.device, memory_format=CL)
            _lib.check(L.rvo_stem_forward(_lib.ptr(buf), Ce, Ci, ks, sd, pd, h, _lib.ptr(ev), _lib.ptr(im),
                                          H, W, _lib.ptr(prev), int(use_image), _lib.ptr(out), st),
                       "rvo_stem_forward")
            self.super_states[k] = out
            per_scale.append(out)
        side = getattr(self, "branch_stream", None)
        _StatArena.begin(ev.device)          # zeroed before the branches fork; both CNNs take slices of it
        if side is None:
            fmap = self.fmap_encoder(per_scale[0], per_scale[1], per_scale[2], out_scale)
            imap = self.imap_encoder(per_scale[0], per_scale[1], per_scale[2], out_scale)
        else:
            # the two CNNs are independent (~45 small kernels each): fork the context encoder onto a side
            # stream so that a captured CUDA graph runs them as parallel branches
            cur = torch.cuda.current_stream(ev.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                imap = self.imap_encoder(per_scale[0], per_scale[1], per_scale[2], out_scale)
            fmap = self.fmap_encoder(per_scale[0], per_scale[1], per_scale[2], out_scale)
            cur.wait_stream(side)
            if not torch.cuda.is_current_stream_capturing():     # eager use: keep the allocator informed
                imap.record_stream(cur)
                for t in per_scale:
                    t.record_stream(side)
        _StatArena.end()
        return fmap[None], imap[None]

    def forward(self, events, images, mask, reinit_hidden=False, out_scale=1.0):
        """events [1,T,Ce,H,W], images [1,Ti,3,H,W], mask [T] bool (one image per True entry)
        -> fmap [1,n,128,H/4,W/4], imap [1,n,384,H/4,W/4] with n = number of images consumed
        (n = 1 with the states of the last event voxel when no image arrived, extractor.py:455)."""
        mask = torch.as_tensor(mask).reshape(-1).tolist()
        if (_fast(events) and events.shape[1] == 1 and len(mask) == 1 and not self.norm_superstate
                and images.shape[1] >= 1):
            return self._forward_fast(events, images, bool(mask[0]), reinit_hidden, out_scale)
        ev = events[0].contiguous(memory_format=torch.channels_last)
        im = images[0].contiguous(memory_format=torch.channels_last)
        per_scale = []
        for k in range(3):
            if reinit_hidden:
                self.super_states[k] = None
            he = self.ev_encoders[k](ev)                 # [T,h,H/s,W/s]
            hi = self.im_encoders[k](im)
            ss = self.super_states[k]
            outs, n_im = [], 0
            for t in range(he.shape[0]):
                ss = self.super_state_ev_encoder[k](he[t:t + 1], ss)
                if mask[t]:
                    ss = self.super_state_im_encoders[k](hi[n_im:n_im + 1], ss)
                    n_im += 1
                    outs.append(ss)
            allss = torch.cat(outs, 0) if outs else ss
            if self.norm_superstate:
                allss = F.instance_norm(allss)
            # the reference carries `norm_super_states[None]` and squeezes it on the next call
            # (extractor.py:440-441,560): with one image per call that is the last state
            self.super_states[k] = allss[-1:] if allss.shape[0] > 1 else allss
            per_scale.append(allss)
        fmap = self.fmap_encoder(per_scale[0], per_scale[1], per_scale[2], out_scale)
        imap = self.imap_encoder(per_scale[0], per_scale[1], per_scale[2], out_scale)
        return fmap[None], imap[None]
