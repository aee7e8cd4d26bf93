"""Event representation on the GPU: the step immediately before the hot path.

`EventToStack` is the device-side twin of utils/transformers.py:128-161 (EventToStack_Numpy, the "stack"
representation every shipped config_net/*.json selects): raw events in arrival order -> the 5-bin int8 event
stack, produced directly as the fp32 [bins,H,W] tensor the RAMP encoder consumes (and as int8 on request), by
rvo_event_stack (csrc/frame_ops.cu) — one scatter pass with exact integer sums + one cast pass, instead of a
host-side `np.add.at` and a 6 MB host->device copy of the finished stack (the raw events are 5 bytes each).
"""
import numpy as np
import torch

from . import _lib


class Events:
    """data/events.py:9-36: x, y (pixel), t (int64), p (int8 polarity; 0 is mapped to -1) + sensor size."""

    def __init__(self, x, y, t, p, width, height):
        self.x, self.y, self.t, self.p = x, y, t, p
        self.width, self.height = width, height
        if len(x):
            self.p[self.p == 0] = -1

    def __len__(self):
        return len(self.x)


class EventToStack:
    """EventToStack_Numpy(num_bins) with a device output.  __call__(events) -> fp32 [bins,H,W] CUDA tensor whose
    values equal `EventToStack_Numpy(num_bins)(events).astype(float32)` bit for bit (integer pixel coordinates,
    the uint16 path of transformers.py:133-135)."""

    def __init__(self, num_bins, device="cuda"):
        self.num_bins = num_bins
        self.device = torch.device(device)

    def __call__(self, events, return_int8=False):
        H, W = int(events.height), int(events.width)
        n = len(events)
        dev = self.device
        out = torch.empty(self.num_bins, H, W, dtype=torch.float32, device=dev)
        out8 = torch.empty(self.num_bins, H, W, dtype=torch.int8, device=dev) if return_int8 else None

        def to_dev(a, np_dtype, t_dtype):
            if torch.is_tensor(a):
                return a.to(device=dev, dtype=t_dtype).contiguous()
            a = np.ascontiguousarray(a)
            if a.dtype != np_dtype:
                if np_dtype == np.uint16 and not np.issubdtype(a.dtype, np.integer):
                    raise RuntimeError("EventToStack: only integer pixel coordinates are supported (the uint16 path "
                                       "of utils/transformers.py:133-135)")
                a = a.astype(np_dtype)
            if np_dtype == np.uint16:       # torch has no uint16 arithmetic; ship the raw 16-bit words
                return torch.from_numpy(a.view(np.int16)).to(dev)
            return torch.from_numpy(a).to(dev)

        x = to_dev(events.x, np.uint16, torch.int16)
        y = to_dev(events.y, np.uint16, torch.int16)
        p = to_dev(events.p, np.float32, torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().rvo_event_stack(_lib.ptr(x), _lib.ptr(y), _lib.ptr(p), n, self.num_bins, H, W,
                                                  _lib.ptr(out), _lib.ptr(out8), _lib.stream_ptr(dev)),
                       "rvo_event_stack")
        return (out, out8) if return_int8 else out
