"""Shared helpers for the parity tests."""
import numpy as np
import torch

from rampvo_b200 import synth


def problem_tensors(prob, device="cuda"):
    """numpy problem (synth.make_problem) -> torch tensors in the reference's layouts."""
    t = {}
    t["poses"] = torch.from_numpy(prob["poses"]).to(device)[None].contiguous()        # [1,n,7]
    t["patches"] = torch.from_numpy(prob["patches"]).to(device)[None].contiguous()    # [1,K,3,P,P]
    t["intrinsics"] = torch.from_numpy(prob["intrinsics"]).to(device)[None].contiguous()
    for k in ("ii", "jj", "kk"):
        t[k] = torch.from_numpy(prob[k]).to(device)
    return t


def targets_from_reprojection(prob, oracle):
    """target = un-clamped reprojection of the patch centre + noise."""
    c = oracle.reproject(prob["poses"], prob["patches"], prob["intrinsics"], prob["ii"], prob["jj"],
                         prob["kk"])
    P = prob["P"]
    ctr = c[:, :, P // 2, P // 2]
    return (ctr + prob["noise"]).astype(np.float32)


def perturb_poses(prob, sigma_t=0.01, sigma_r=0.005, seed=7):
    """Perturbs the free poses so that BA has something to do."""
    rng = np.random.default_rng(seed)
    p = prob["poses"].astype(np.float64).copy()
    for f in range(prob["t0"], prob["t1"]):
        p[f, :3] += rng.normal(0, sigma_t, 3)
        p[f, 3:] += rng.normal(0, sigma_r, 4)
        p[f, 3:] /= np.linalg.norm(p[f, 3:])
    return p.astype(np.float32)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)
