"""Seeded synthetic inputs shaped like the reference's steady-state VO.

The patch graph is produced by replaying the edge rules of ramp/Ramp_vo.py:312-325 (forward /
backward edges when a frame is added) and :273 (edges older than REMOVAL_WINDOW are dropped), with
no keyframe removals (upper bound on the graph size).  numpy only; used by tests and bench.py.
"""
import numpy as np

CONFIGS = {
    # name: (PATCHES_PER_FRAME, PATCH_LIFETIME, REMOVAL_WINDOW, OPTIMIZATION_WINDOW)
    "cfg1": (32, 13, 22, 10),      # BASELINE.json configs[0]: 32 patches, 8 frames
    "fast": (48, 11, 16, 7),       # config_vo/fast.yaml
    "default": (96, 13, 22, 10),   # config_vo/default.yaml
    "precise": (300, 33, 42, 30),  # config_vo/precise.yaml
}


def replay_graph(M, lifetime, removal, n_frames):
    """Returns (ii, jj, kk) int64 after `n_frames` frames have been added (Ramp_vo.__call__ order:
    append forward edges, append backward edges, update, drop old edges)."""
    ii = np.zeros(0, np.int64)
    jj = np.zeros(0, np.int64)
    kk = np.zeros(0, np.int64)
    for n in range(1, n_frames + 1):         # n = number of frames after adding frame n-1
        m = n * M
        r = lifetime
        # __edges_forw (Ramp_vo.py:312-318): patches of frames [max(n-r,0), n-1) -> frame n-1
        t0 = M * max(n - r, 0)
        t1 = M * max(n - 1, 0)
        kf, jf = np.meshgrid(np.arange(t0, t1), np.arange(n - 1, n), indexing="ij")
        # __edges_back (Ramp_vo.py:320-325): patches of frame n-1 -> frames [max(n-r,0), n)
        kb, jb = np.meshgrid(np.arange(max(m - M, 0), m), np.arange(max(n - r, 0), n), indexing="ij")
        knew = np.concatenate([kf.reshape(-1), kb.reshape(-1)])
        jnew = np.concatenate([jf.reshape(-1), jb.reshape(-1)])
        jj = np.concatenate([jj, jnew])
        kk = np.concatenate([kk, knew])
        ii = np.concatenate([ii, knew // M])
        # remove_factors(ii < n - REMOVAL_WINDOW) (Ramp_vo.py:273)
        keep = ~(ii < n - removal)
        ii, jj, kk = ii[keep], jj[keep], kk[keep]
    return ii, jj, kk


def make_problem(config="default", n_frames=40, seed=0, ht=120, wd=160, P=3, noise_px=1.0,
                 n_pose_slots=None):
    """Op-level problem: graph + poses + patches + intrinsics + targets/weights.
    Returns a dict of numpy arrays (float32 / int64)."""
    M, lifetime, removal, optwin = CONFIGS[config]
    rng = np.random.default_rng(seed)
    ii, jj, kk = replay_graph(M, lifetime, removal, n_frames)
    n = n_frames
    slots = n_pose_slots or n
    # poses: chain of small motions, pose 0 = identity
    from numpy.linalg import norm
    poses = np.zeros((slots, 7), np.float64)
    poses[:, 6] = 1.0
    t = np.zeros(3)
    q = np.array([0, 0, 0, 1.0])
    for f in range(1, n):
        xi = rng.normal(0, 0.02, 6)
        th = norm(xi[3:])
        dq = np.concatenate([np.sin(th / 2) * xi[3:] / max(th, 1e-12), [np.cos(th / 2)]])
        # left-multiply: T <- Exp(xi) T
        x, y, z, w = dq
        X, Y, Z, W = q
        qn = np.array([w * X + x * W + y * Z - z * Y, w * Y - x * Z + y * W + z * X,
                       w * Z + x * Y - y * X + z * W, w * W - x * X - y * Y - z * Z])
        uv = 2 * np.cross(dq[:3], t)
        t = t + dq[3] * uv + np.cross(dq[:3], uv) + xi[:3]
        q = qn / norm(qn)
        poses[f, :3], poses[f, 3:] = t, q
    K = n * M
    cx = rng.uniform(8, wd - 8, K)
    cy = rng.uniform(8, ht - 8, K)
    d = rng.uniform(0.2, 2.0, K)
    g = np.arange(P) - P // 2
    patches = np.zeros((K, 3, P, P), np.float64)
    patches[:, 0] = cx[:, None, None] + g[None, None, :]
    patches[:, 1] = cy[:, None, None] + g[None, :, None]
    patches[:, 2] = d[:, None, None]
    intr = np.tile(np.array([wd * 0.5, wd * 0.5, wd * 0.5, ht * 0.5]), (slots, 1))  # 320/4 .. (evaluate.py:48 / RES)
    E = len(ii)
    out = dict(ii=ii, jj=jj, kk=kk, poses=poses.astype(np.float32),
               patches=patches.astype(np.float32), intrinsics=intr.astype(np.float32),
               M=M, n=n, t0=max(n - optwin, 1), t1=n, P=P, ht=ht, wd=wd, E=E)
    out["noise"] = rng.normal(0, noise_px, (E, 2)).astype(np.float32)
    out["weight"] = rng.uniform(0, 1, (E, 2)).astype(np.float32)
    return out


def make_features(n_frames_ring, n_patches_ring, C=128, ht=120, wd=160, P=3, seed=0,
                  dtype=np.float16, levels=(1, 4)):
    """gmap [Np,P,P,C] and pyramid levels [Nf,H,W,C] (channels-last), N(0,1)/sqrt(C)."""
    rng = np.random.default_rng(seed + 1000)
    s = 1.0 / np.sqrt(C)
    gmap = (rng.standard_normal((n_patches_ring, P, P, C), dtype=np.float32) * s).astype(dtype)
    pyr = [(rng.standard_normal((n_frames_ring, ht // l, wd // l, C), dtype=np.float32) * s).astype(dtype)
           for l in levels]
    return gmap, pyr


class SyntheticSequence:
    """TartanEvent-shaped synthetic stream: per call one 5-bin event stack
    [1,1,5,H,W] (integer-valued fp32, ~10 % non-zero, utils/transformers.py:149-161) and one image
    [1,1,3,H,W] in [-0.5,1.5] (ramp/utils.py:582): a smooth random texture translated a few px per
    frame, with the event rate following the texture gradient so top-k/NMS spreads the patches."""

    def __init__(self, ht=480, wd=640, seed=0, device="cpu", shift=(3.0, 1.5)):
        import torch
        import torch.nn.functional as F
        self.torch, self.F = torch, F
        self.ht, self.wd, self.device = ht, wd, device
        self.shift = shift
        g = torch.Generator().manual_seed(seed)
        pad = 256
        low = torch.rand(1, 3, (ht + 2 * pad) // 16, (wd + 2 * pad) // 16, generator=g)
        self.canvas = F.interpolate(low, size=(ht + 2 * pad, wd + 2 * pad), mode="bicubic",
                                    align_corners=False).clamp(0, 1).to(device)
        self.pad = pad
        self.gen = torch.Generator(device=device).manual_seed(seed + 1)
        self.intrinsics = torch.tensor([320.0, 320.0, 320.0, 240.0])   # evaluate.py:48

    def frame(self, t):
        torch, F = self.torch, self.F
        dx, dy = self.shift[0] * t, self.shift[1] * t
        x0 = int(self.pad + (dx % self.pad) * (1 if (int(dx) // self.pad) % 2 == 0 else -1))
        y0 = int(self.pad + (dy % self.pad) * (1 if (int(dy) // self.pad) % 2 == 0 else -1))
        x0 = max(0, min(x0, self.canvas.shape[-1] - self.wd))
        y0 = max(0, min(y0, self.canvas.shape[-2] - self.ht))
        img = self.canvas[:, :, y0:y0 + self.ht, x0:x0 + self.wd]
        image = (2.0 * img - 0.5)[None]                                   # [1,1,3,H,W]
        gray = img.mean(1, keepdim=True)
        gx = F.pad(gray[..., :, 1:] - gray[..., :, :-1], (0, 1))
        gy = F.pad(gray[..., 1:, :] - gray[..., :-1, :], (0, 0, 0, 1))
        rate = (gx.abs() + gy.abs())
        rate = 0.12 * rate / rate.mean().clamp(min=1e-6)
        lam = rate.expand(1, 5, self.ht, self.wd).contiguous()
        ev = torch.poisson(lam, generator=self.gen)
        sign = torch.randint(0, 2, ev.shape, generator=self.gen, device=self.device) * 2 - 1
        events = (ev * sign).clamp(-127, 127)[None]                        # [1,1,5,H,W]
        mask = torch.tensor([True])
        return events, image, mask
