"""Drop-in for ramp/pose_prediction/pose_pred_utils.py — the optional pose-prediction branch of Ramp_vo
(`use_pose_pred`, ramp/Ramp_vo.py:412-545; off in every shipped config): extrapolate each patch track onto a virtual
future frame with a per-patch smoothing spline, then let fastba place the virtual frame.

The reference runs this on the host, one patch at a time, with a device round trip per patch per call
(`.cpu().item()` in compute_patch_track__, :171-187).  Same algorithm and the same scipy spline here, but the device is
read ONCE per call: the tracks of every patch are gathered with one sort + one copy, the models are fitted on the host
(scipy.interpolate.UnivariateSpline, as the reference: pose_pred_utils.py:271-287) and the predicted targets are
written back with one indexed store.
"""
from collections import OrderedDict

import numpy as np
import torch

from .vo_utils import flatmeshgrid

PAST_PATCH_NUM = 5          # pose_pred_utils.py:262: the last 5 observations of a track feed the spline


def motion_bootstrap(n, poses, MOTION_MODEL, MOTION_DAMPING):
    """pose_pred_utils.py:190-199: damped-linear extrapolation of the last two poses ([N,7] tensor)"""
    from .net import motion_bootstrap as mb          # one implementation (ramp/net.py imports it from here too)
    return mb(n=n, poses=poses, MOTION_MODEL=MOTION_MODEL, MOTION_DAMPING=MOTION_DAMPING)


def add_forward_elements(frame_num, patch_extracted_num, r, ii, jj, kk, ix, weights):
    """pose_pred_utils.py:202-215: edges from every patch of the last r-1 frames to the virtual frame
    `frame_num - 1`; their confidences start at zero"""
    dev = ii.device
    t0 = patch_extracted_num * max(frame_num - r, 0)
    t1 = patch_extracted_num * max(frame_num - 1, 0)
    kk_add, jj_add = flatmeshgrid(torch.arange(t0, t1, device=dev), torch.arange(frame_num - 1, frame_num, device=dev),
                                  indexing='ij')
    ii_s = torch.cat([ii, ix[kk_add]])
    jj_s = torch.cat([jj, jj_add])
    kk_s = torch.cat([kk, kk_add])
    w_s = torch.cat([weights, torch.zeros((1, len(kk_add), 2), device=dev, dtype=weights.dtype)], dim=1)
    return ii_s, jj_s, kk_s, w_s


def compute_patch_track(coords, ii, jj, kk, image_to_proj):
    """compute_patch_track__ (pose_pred_utils.py:171-187): for every (source frame, patch) that has an edge to frame
    `image_to_proj`, the [n_obs, 2] track of pixel (0, 0) of the patch over ALL its edges, in edge order.
    Returns an OrderedDict keyed like the reference's (first-seen order of the new edges), numpy values."""
    ii_h, jj_h, kk_h = ii.cpu().numpy(), jj.cpu().numpy(), kk.cpu().numpy()
    xy = coords[0, :, :, 0, 0].detach().float().cpu().numpy()                  # one device read
    new = np.nonzero(jj_h == image_to_proj)[0]
    order = np.argsort(kk_h, kind="stable")                                    # edges of a patch, in edge order
    ks = kk_h[order]
    tracks = OrderedDict()
    for e in new:
        key = (int(ii_h[e]), int(kk_h[e]))
        if key in tracks:
            continue
        lo, hi = np.searchsorted(ks, key[1], "left"), np.searchsorted(ks, key[1], "right")
        idx = order[lo:hi]
        idx = idx[ii_h[idx] == key[0]]
        if len(idx):
            tracks[key] = xy[idx]
    return tracks


def fit_model_patch_track(next_frame_index, patch_dict, img_to_keyframe_map, ii, jj, data_shape, frequency=30, deg=2):
    """pose_pred_utils.py:245-291: one weighted smoothing spline per coordinate per track over its last 5
    observations (the edge to the virtual frame, last in the track, is dropped); a track whose last 5 observations
    all left the image gets confidence 0, the others 1e-9."""
    from scipy.interpolate import UnivariateSpline
    height, width = data_shape
    ii_h = ii.cpu().numpy() if torch.is_tensor(ii) else np.asarray(ii)
    jj_h = jj.cpu().numpy() if torch.is_tensor(jj) else np.asarray(jj)
    tmap = (img_to_keyframe_map.cpu().numpy() if torch.is_tensor(img_to_keyframe_map)
            else np.asarray(img_to_keyframe_map))
    first_of = {}
    models = OrderedDict()
    for key, track in patch_dict.items():
        start_image = key[0]
        if start_image not in first_of:
            first_of[start_image] = int(jj_h[ii_h == start_image].min())
        track = np.asarray(track.cpu() if torch.is_tensor(track) else track)       # float32, as the reference feeds scipy
        x, y = track[:-1].T
        # torch divides the int64 timestamps in float32 (pose_pred_utils.py:260)
        t = tmap[first_of[start_image]:next_frame_index].astype(np.float32) / np.float32(frequency)
        mask = (x >= 0) & (x < width) & (y >= 0) & (y < height)
        masked_weights = 0 if np.all(mask[-PAST_PATCH_NUM:] == False) else 10 ** -9     # noqa: E712
        x_, y_, t_ = x[-PAST_PATCH_NUM:], y[-PAST_PATCH_NUM:], t[-PAST_PATCH_NUM:]
        w = (t_ - t_[0]) / (t[-1] - t_[0]) + 10 ** -7
        assert len(t_) == len(x_)
        spl_x = UnivariateSpline(x=t_, y=x_, w=w, bbox=[None, None], k=deg, s=None, ext=0, check_finite=False)
        spl_y = UnivariateSpline(x=t_, y=y_, w=w, bbox=[None, None], k=deg, s=None, ext=0, check_finite=False)
        models[key] = (spl_x, spl_y, masked_weights, t_[-1])
    return models


def predict_patch_on_model(patch_models, step_to_pred_future, frequency, next_frame_index, coords, weights, ii, jj, kk):
    """pose_pred_utils.py:294-319: evaluate every track's splines `step_to_pred_future` frames ahead and write the
    3x3 grid around the prediction into the virtual-frame edge of that patch (coords [1,E,2,3,3], weights [1,E,2]).
    One indexed store instead of a masked store per patch."""
    if not patch_models:
        return coords, weights
    dev = coords.device
    ii_h, jj_h, kk_h = ii.cpu().numpy(), jj.cpu().numpy(), kk.cpu().numpy()
    virt = np.nonzero(jj_h == next_frame_index)[0]
    lut = {}
    for e in virt:
        lut.setdefault((int(ii_h[e]), int(kk_h[e])), []).append(int(e))
    rows, grids, ws = [], [], []
    off = np.arange(-1.0, 2.0)
    for key, (spl_x, spl_y, masked_weights, last_t) in patch_models.items():
        es = lut.get(key)
        if not es:
            continue
        new_time = last_t + step_to_pred_future / frequency
        nx, ny = float(spl_x(new_time)), float(spl_y(new_time))
        # the 3x3 grid in the layout of Ramp_vo.reproject: channel 0 = x (varies along the last axis), channel 1 = y
        # (varies along the row axis).  The reference stacks (rows_grid, cols_grid) of torch.meshgrid(x, y), which
        # puts the y prediction into the x channel (pose_pred_utils.py:309-316) — not reproduced.
        g = np.stack([np.repeat((nx + off)[None, :], 3, axis=0), np.repeat((ny + off)[:, None], 3, axis=1)], 0)
        for e in es:
            rows.append(e)
            grids.append(g)
            ws.append(masked_weights)
    if rows:
        idx = torch.as_tensor(rows, device=dev, dtype=torch.long)
        coords[0, idx] = torch.as_tensor(np.stack(grids), device=dev, dtype=coords.dtype)
        weights[0, idx] = torch.as_tensor(ws, device=dev, dtype=weights.dtype)[:, None].expand(-1, 2)
    return coords, weights
