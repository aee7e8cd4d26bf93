"""ctypes binding of librampvo_b200.so — the C-ABI declared in include/rampvo_b200.h.

This is the ONLY way the Python host side reaches the CUDA kernels: plain pointers and sizes, no
torch types cross the boundary.  There is no fallback: if the library is missing or a call fails,
a RuntimeError is raised (the reference surfaces C++ exceptions as RuntimeError too).
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librampvo_b200.so")

RVO_F16, RVO_F32 = 0, 1
RVO_TF_TONLY, RVO_TF_NOCLAMP = 1, 2


class FMap(Structure):
    """rvo_fmap_t"""
    _fields_ = [("data", c_void_p), ("dtype", c_int32), ("N", c_int32), ("C", c_int32),
                ("H", c_int32), ("W", c_int32), ("sN", c_int64), ("sC", c_int64),
                ("sH", c_int64), ("sW", c_int64)]


class ChainLayer(Structure):
    """rvo_chain_layer_t"""
    _fields_ = [("w16", c_void_p), ("bias16", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
                ("y16", c_void_p), ("ldy", c_int64), ("K", c_int32), ("epilogue", c_int32)]


class Chain(Structure):
    """rvo_chain_t (include/rampvo_b200.h): one row-local stretch of the update operator"""
    _fields_ = [("M", c_int32), ("n_layers", c_int32), ("prologue", c_int32), ("reserved", c_int32),
                ("a16", c_void_p), ("lda", c_int64), ("gather", c_void_p),
                ("x32", c_void_p), ("hy_a", c_void_p), ("grp_a", c_void_p), ("hy_b", c_void_p), ("grp_b", c_void_p),
                ("pro_gamma", c_void_p), ("pro_beta", c_void_p),
                ("net_in", c_void_p), ("imap16", c_void_p), ("imap_idx", c_void_p), ("imap_mod", c_int64),
                ("res32", c_void_p), ("out32", c_void_p), ("out16", c_void_p),
                ("Wd", c_void_p), ("bd", c_void_p), ("Ww", c_void_p), ("bw", c_void_p),
                ("delta", c_void_p), ("weight", c_void_p),
                ("scratch32", c_void_p), ("scratch16", c_void_p),
                ("layer", ChainLayer * 6)]


PRO_ROWS, PRO_EXPAND, PRO_EXPAND_LN = 0, 1, 2
(EPI_RELU, EPI_LN_RELU, EPI_ADD3_LN, EPI_RES, EPI_STORE16, EPI_GATE, EPI_GATED_LN, EPI_GATED_HEADS) = range(8)

_P = c_void_p
_I64 = c_int64
_SIGNATURES = {
    "rvo_up_chain_scratch_rows": (c_int64, []),
    "rvo_up_chain": (c_int, [POINTER(Chain), c_void_p]),
    "rvo_up_chain_info": (c_int, [POINTER(c_int), POINTER(c_int)]),
    "rvo_up_chain_set_cluster": (c_int, [c_int]),
    "rvo_plan_edge_groups": (c_int, [c_void_p, c_int, POINTER(c_void_p)]),
    "rvo_abi_version": (c_int, []),
    "rvo_last_error": (c_char_p, []),
    "rvo_device_cc": (c_int, []),
    "rvo_launch_count": (ctypes.c_uint64, []),
    "rvo_set_sm_budget": (c_int, [c_int]),
    "rvo_net_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "rvo_edges_step_tiles": (c_int64, [c_int]),
    "rvo_edges_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_int, c_void_p, c_void_p, c_void_p, ctypes.c_uint32, c_void_p, c_int, c_void_p,
                               c_void_p]),
    "rvo_get_sm_budget": (c_int, []),
    "rvo_patchify_forward": (c_int, [POINTER(FMap), _P, c_int, c_int, _P, _P]),
    "rvo_patchify_bilinear": (c_int, [POINTER(FMap), _P, c_int, c_int, _P, c_int, _I64, _I64, _I64,
                                      _I64, _I64, _P]),
    "rvo_corr_forward": (c_int, [POINTER(FMap), POINTER(FMap), _P, _P, _P, c_int, c_int, _P, _P]),
    "rvo_corr_pyramid": (c_int, [POINTER(FMap), POINTER(FMap), POINTER(c_float), c_int, _P, _P, _P,
                                 _I64, _I64, c_int, c_int, _P, _I64, _P]),
    "rvo_corr_backward": (c_int, [POINTER(FMap), POINTER(FMap), _P, _P, _P, _P, c_int, c_int, _P, _P, _P]),
    "rvo_patchify_backward": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "rvo_up_linear": (c_int, [_P, _I64, _P, _P, c_int, c_int, c_int, c_int, _P, _I64, _P]),
    "rvo_up_linear_gather": (c_int, [_P, _I64, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _I64, _P]),
    "rvo_corr_tiles_ws_bytes": (_I64, [POINTER(FMap), c_int, c_int]),
    "rvo_corr_tiles": (c_int, [POINTER(FMap), POINTER(FMap), POINTER(c_float), c_int, _P, _P, _P, _I64,
                               _I64, c_int, _P, _I64, _P, _I64, _P]),
    "rvo_corr_pyramid_host": (c_int, [POINTER(FMap), POINTER(FMap), POINTER(c_float), c_int, _P, _P,
                                      _P, _I64, _I64, c_int, c_int, _P, _P]),
    "rvo_transform": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "rvo_transform_jac": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "rvo_reproject": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, _P, _P]),
    "rvo_point_cloud": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P]),
    "rvo_flow_mag": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_float, _P, _P]),
    "rvo_in_stats": (c_int, [_P, _I64, c_int, _P, _P]),
    "rvo_in_apply": (c_int, [_P, _P, _P, _P, _I64, c_int, c_float, _P, _P]),
    "rvo_stem_params_layout": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "rvo_stem_forward": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int, _P,
                                 c_int, _P, _P]),
    "rvo_motion_model": (c_int, [_P, c_int, c_float, _P]),
    "rvo_pair_flow": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, _I64, _I64, c_float, _P, _P]),
    "rvo_plan_bytes": (_I64, [c_int]),
    "rvo_graph_plan": (c_int, [_P, _P, c_int, _I64, _I64, _P, _I64, _P]),
    "rvo_plan_groups": (c_int, [_P, c_int, POINTER(_P), POINTER(_P), POINTER(_P), POINTER(_P),
                                POINTER(_P)]),
    "rvo_neighbors_ws_bytes": (_I64, [c_int]),
    "rvo_neighbors": (c_int, [_P, _P, c_int, _I64, _I64, _P, _P, _P, _I64, _P]),
    "rvo_plan_neighbors": (c_int, [_P, c_int, _P, _P, _P]),
    "rvo_ba_ws_bytes": (_I64, [c_int, _I64, c_int]),
    "rvo_ba_forward": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _I64, _I64, c_int, c_int,
                               c_int, c_int, c_int, c_int, _P, _I64, _P]),
    "rvo_ba_forward_planned": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _I64, _I64, c_int,
                                       c_int, c_int, c_int, _P, _I64, _P]),
    "rvo_ba_forward_dyn": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _I64, _I64, c_int, c_int,
                                   _P, c_int, _P, _I64, _P]),
    "rvo_ba_forward_fused": (c_int, [_P, _P, _P, _P, _P, _P, c_float, c_float, _P, _P, _P, _P, _P, c_int, _I64,
                                     c_int, c_int, _P, c_int, _P, _I64, _P]),
    "rvo_select_ws_bytes": (_I64, [c_int, c_int]),
    "rvo_select_patches": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _I64, _P]),
    "rvo_pyramid_level2": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "rvo_copy_segments": (c_int, [POINTER(_P), POINTER(_P), POINTER(_I64), c_int, _P]),
    "rvo_conv2d_kpad": (c_int, [c_int, c_int]),
    "rvo_conv2d_nhwc": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P, _P]),
    "rvo_scene_lstm_params_floats": (c_int, [c_int, c_int]),
    "rvo_scene_lstm_forward": (c_int, [_P, c_int, c_int, _P, _P, c_int, c_int, _P, _P, _P, _P, c_int, _P, _P]),
    "rvo_frame_commit": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, POINTER(c_float), c_int, c_int, c_int, c_int,
                                 _I64, _I64, c_int, _P]),
    "rvo_event_stack": (c_int, [_P, _P, _P, _I64, c_int, c_int, c_int, _P, _P, _P]),
    "rvo_ba_plan": (c_int, [_P, _P, c_int, _I64, _I64, c_int, _P, _I64, _P]),
    "rvo_ba_assemble": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, _I64, c_int, c_int, c_int,
                                _P, _P, _I64, _P]),
    "rvo_ba_assemble_fused": (c_int, [_P, _P, _P, _P, _P, _P, c_float, c_float, _P, _P, _P, _P, c_int, _I64, c_int,
                                      c_int, c_int, _P, _P, _I64, _P]),
    "rvo_ba_solve": (c_int, [_P, _P, _P, c_int, _I64, c_int, c_int, c_int, _P, _I64, _P]),
    "rvo_softagg": (c_int, [_P, _P, c_int, _P, c_int, c_int, _I64, _P, c_int, _P]),
    "rvo_expand_add": (c_int, [_P, c_int, _P, c_int, c_int, _P, _P]),
    "rvo_gather_rows": (c_int, [_P, _P, c_int, c_int, _P, c_int, _P]),
    "rvo_up_ln_relu": (c_int, [_P, _P, _P, c_int, c_int, _P, _P]),
    "rvo_up_add3_ln": (c_int, [_P, _P, _P, _I64, _P, _P, _P, c_int, c_int, _P, _P, _P]),
    "rvo_up_add_cast": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "rvo_up_softagg_fg": (c_int, [_P, _P, c_int, c_int, _I64, _P, _P]),
    "rvo_up_expand_add_ln": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "rvo_up_gated_tail": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                  _P, _P]),
    "rvo_ba_solve_poses": (c_int, [_P, _P, c_int, c_int, _P, _I64, _P]),
    "rvo_ba_forward_host": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _I64, _I64, c_int,
                                    c_int, c_int, c_int, c_int, c_int, _P]),
}

_lib = None


def exported_symbols():
    """Names include/rampvo_b200.h declares (used by the CPU test that checks the exports)."""
    return sorted(_SIGNATURES)


def lib():
    """Loads the library once; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "librampvo_b200.so is missing (%s): run `python -m rampvo_b200.build`; "
                "there is no CPU / PyTorch fallback for the hot path" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.rvo_abi_version() != 1:
            raise RuntimeError("librampvo_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().rvo_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device (or host) pointer of a tensor, or NULL for None."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("rampvo_b200: expected a CUDA tensor (the hot path has no CPU fallback)")


def dtype_code(t):
    if t.dtype == torch.float16:
        return RVO_F16
    if t.dtype == torch.float32:
        return RVO_F32
    raise RuntimeError("rampvo_b200: unsupported feature dtype %s (float16 / float32 only)" % t.dtype)


def fmap_view(t):
    """rvo_fmap_t for a 4-D tensor view [N,C,H,W] with arbitrary strides."""
    assert t.dim() == 4
    s = t.stride()
    return FMap(t.data_ptr(), dtype_code(t), t.shape[0], t.shape[1], t.shape[2], t.shape[3],
                s[0], s[1], s[2], s[3])


class Workspace:
    """Grow-only device scratch buffers, one per (device, tag, stream): two Ramp_vo instances on different streams
    never share scratch.  A CUDA graph that captured a buffer must keep it alive: `snapshot()` returns the tensors
    currently registered for a device so that the graph object can hold references (a buffer that later grows is
    replaced here, but the captured one stays valid for as long as its graph lives)."""
    _bufs = {}

    @classmethod
    def get(cls, device, nbytes, tag="default"):
        dev = torch.device(device)
        key = (str(dev), tag, torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else 0)
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
            cls._bufs[key] = buf
        return buf

    @classmethod
    def snapshot(cls, device):
        d = str(torch.device(device))
        return [b for (dev, _, _), b in cls._bufs.items() if dev == d]
