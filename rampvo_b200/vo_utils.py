"""The hot-path helpers of ramp/utils.py (the rest of that file — losses, plotting, pose-file I/O —
is out of scope).  Small tensor plumbing; stays in PyTorch."""
import torch
import torch.nn.functional as F

from . import _lib


def get_channel_dim(cfg):
    """ramp/utils.py:244-245"""
    return (cfg["num_event_bins"], 3)


def check_input_tensors(events, images):
    """ramp/utils.py:229-241: 5-D [batch, n, C, H, W], batch 1."""
    if events.dim() != images.dim():
        raise AssertionError("Event and image tensor must have the same number of dimension")
    if events.dim() != 5:
        raise AssertionError("Event and image tensor must have shape [batch, n_tensors, channels, height, width]")
    if not (events.shape[0] == 1 and images.shape[0] == 1):
        raise NotImplementedError("Event and image tensor must have batch dimension (0 dim) = 1")


def preprocess_input(input_tensor):
    """ramp/utils.py:250-256 (a 2-tuple leaves `mask` undefined there; it is an error here too)."""
    if len(input_tensor) != 3:
        raise ValueError("input_tensor must be (events, images, mask)")
    events, images, mask = input_tensor
    check_input_tensors(events, images)
    return events, images, mask


def nms_image(x, kernel_size=3):
    """ramp/utils.py:157-183: keep values equal to their (k x k) neighbourhood maximum."""
    mx = F.max_pool2d(x[None], kernel_size, stride=1, padding=(kernel_size - 1) // 2)[0]
    return x * (mx == x).float()


def coords_from_topk_events(events, patches_per_image, border_suppression_size=0, non_max_supp_rad=0):
    """ramp/utils.py:186-226 (get_coords_from_topk_events).  events [1,n,C,H,W] -> coords [n,M,2].
    The map is transposed to (x, y) order and the flat top-k index is split with a TRUE division
    (utils.py:212), so x = x_int + y/H' is fractional while y is integral — preserved as is."""
    if events.is_cuda:
        return _select_patches_cuda(events, patches_per_image, border_suppression_size, non_max_supp_rad)
    # host tensors (CPU tests of the selection rule against the reference fixture): the reference's own torch ops
    ev = torch.abs(events.squeeze(0))
    ev = F.avg_pool2d(ev, 4, 4).transpose(3, 2)
    ev_mean = torch.mean(ev, dim=1)                      # [n, W', H']
    b = border_suppression_size
    if b != 0:
        ev_mean[:, :b, :] = 0
        ev_mean[:, -b:, :] = 0
        ev_mean[:, :, :b] = 0
        ev_mean[:, :, -b:] = 0
    if non_max_supp_rad != 0:
        ev_mean = nms_image(ev_mean, kernel_size=non_max_supp_rad)
    flat = torch.flatten(ev_mean, start_dim=1)
    _, indices = torch.topk(flat, k=patches_per_image, dim=-1)
    rows = indices / ev_mean.shape[-1]
    cols = indices % ev_mean.shape[-1]
    return torch.stack((rows, cols.to(rows.dtype)), dim=-1)


def _select_patches_cuda(events, M, border, nms):
    """rvo_select_patches (csrc/frame_ops.cu): scoring pass + one CTA doing NMS and an exact top-M, instead of
    ~14 torch launches (the top-k alone is a 108 us single-CTA kernel).  events [1,n,C,H,W] fp32 on the device."""
    L = _lib.lib()
    ev = events.squeeze(0)
    if ev.dtype != torch.float32 or not ev.is_contiguous():
        ev = ev.float().contiguous()
    n, C, H, W = ev.shape
    coords = torch.empty(n, M, 2, dtype=torch.float32, device=ev.device)
    nb = L.rvo_select_ws_bytes(H, W)
    ws = _lib.Workspace.get(ev.device, max(nb, 16), "select")
    # torch.topk on CUDA ends with a key/value sort that is stable for k > 32 (ties by ascending index — done in
    # the kernel) and an unstable bitonic network for k <= 32: there the kernel returns the pre-sort order and the
    # very same torch sort is applied, so ties come out exactly as the reference's torch.topk orders them
    small = M <= 32
    idx = torch.empty(n, M, dtype=torch.int64, device=ev.device) if small else None
    val = torch.empty(n, M, dtype=torch.float32, device=ev.device) if small else None
    with torch.cuda.device(ev.device):
        for f in range(n):
            _lib.check(L.rvo_select_patches(_lib.ptr(ev[f]), C, H, W, M, int(border), int(nms), int(small),
                                            _lib.ptr(coords[f]), _lib.ptr(idx[f]) if small else None,
                                            _lib.ptr(val[f]) if small else None, _lib.ptr(ws), ws.numel(),
                                            _lib.stream_ptr(ev.device)), "rvo_select_patches")
    if small:
        _, pos = torch.sort(val, dim=-1, descending=True)
        idx = torch.gather(idx, 1, pos)
        coords = torch.stack((idx / (H // 4), (idx % (H // 4)).to(torch.float32)), dim=-1)
    return coords


def pyramid_level2(f_cl):
    """F.avg_pool2d(fmap, 4, 4) (ramp/Ramp_vo.py:381) on a channels-last map [H,W,C] -> [H/4,W/4,C]"""
    H, W, C = f_cl.shape
    out = torch.empty(H // 4, W // 4, C, dtype=f_cl.dtype, device=f_cl.device)
    with torch.cuda.device(f_cl.device):
        _lib.check(_lib.lib().rvo_pyramid_level2(_lib.ptr(f_cl), _lib.dtype_code(f_cl), H, W, C, _lib.ptr(out),
                                                 _lib.stream_ptr(f_cl.device)), "rvo_pyramid_level2")
    return out


def copy_segments(pairs):
    """[(src, dst), ...] contiguous same-size device tensors -> one launch (rvo_copy_segments)"""
    import ctypes
    n = len(pairs)
    assert 0 < n <= 8
    src = (ctypes.c_void_p * n)(*[p[0].data_ptr() for p in pairs])
    dst = (ctypes.c_void_p * n)(*[p[1].data_ptr() for p in pairs])
    nb = (ctypes.c_int64 * n)(*[p[0].numel() * p[0].element_size() for p in pairs])
    for a, b in pairs:
        if not (a.is_contiguous() and b.is_contiguous() and a.dtype == b.dtype and a.numel() == b.numel()):
            raise RuntimeError("copy_segments: segments must be contiguous and of equal size / dtype")
    dev = pairs[0][0].device
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().rvo_copy_segments(src, dst, nb, n, _lib.stream_ptr(dev)), "rvo_copy_segments")


def flatmeshgrid(*args, **kwargs):
    """ramp/utils.py:104-106"""
    return (x.reshape(-1) for x in torch.meshgrid(*args, **kwargs))


def filter_features(confidences, target, data_shape):
    """ramp/utils.py:557-570: zero the weights of targets outside [0,wd] x [0,ht]."""
    ht, wd = data_shape
    x, y = target[..., 0], target[..., 1]
    bad = (x < 0) | (x > wd) | (y < 0) | (y > ht)
    return confidences * (~bad)[..., None].to(confidences.dtype)
