"""`torch_scatter` stand-in (TEST / BASELINE INFRASTRUCTURE): the two functions the reference calls
(ramp/blocks.py:44-45, ramp/ba.py:42-56) written with index ops; works on CPU and CUDA tensors.
The real package is unpinned in the reference's requirements.txt:4 and absent from this image."""
import torch


def scatter_sum(src, index, dim=1, dim_size=None):
    assert dim == 1
    n = int(index.max()) + 1 if dim_size is None else dim_size
    out = torch.zeros((src.shape[0], n) + tuple(src.shape[2:]), dtype=src.dtype, device=src.device)
    with torch.autocast(src.device.type, enabled=False):
        return out.index_add_(1, index, src)


def scatter_softmax(src, index, dim=1):
    assert dim == 1
    n = int(index.max()) + 1
    idx = index.view(1, -1, *([1] * (src.dim() - 2))).expand_as(src)
    mx = torch.full((src.shape[0], n) + tuple(src.shape[2:]), -float("inf"), dtype=src.dtype, device=src.device)
    # torch_scatter computes in the dtype of `src` (custom ops are untouched by autocast; its softmax uses the
    # in-place `exp_()` and a scatter_sum in that dtype): fp16 in, fp16 arithmetic, fp16 out under mixed precision
    with torch.autocast(src.device.type, enabled=False):
        mx = mx.scatter_reduce(1, idx, src, reduce="amax", include_self=True)
        ex = (src - mx[:, index]).exp_()
        den = torch.zeros_like(mx).index_add_(1, index, ex)
        return ex.div(den[:, index])
