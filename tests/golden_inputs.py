"""Seeded inputs shared by tests/golden/make_golden.py (which feeds them to the reference's Python
code) and by the tests (which feed them to the oracle / the CUDA path)."""
import numpy as np
import torch

from rampvo_b200 import synth
from tests.util import perturb_poses, targets_from_reprojection

UPDATE_SEED = 1234      # evaluate.py:40 seeds everything with 1234
ENCODER_SEED = 1234


def as_torch(prob, dtype=torch.float32, device="cpu"):
    t = {k: torch.from_numpy(prob[k]).to(device) for k in ("ii", "jj", "kk")}
    t["poses"] = torch.from_numpy(prob["poses"]).to(device=device, dtype=dtype)[None]
    t["patches"] = torch.from_numpy(prob["patches"]).to(device=device, dtype=dtype)[None]
    t["intrinsics"] = torch.from_numpy(prob["intrinsics"]).to(device=device, dtype=dtype)[None]
    return t


def pops_problem():
    return synth.make_problem("cfg1", 8, seed=31)


def ba_problem(oracle):
    """Benign graph (every edge in bounds, poses 0..3 fixed) so that ramp/ba.py and cuda_ba agree."""
    prob = synth.make_problem("cfg1", 8, seed=32, noise_px=0.5)
    prob["t0"] = 4
    tgt = targets_from_reprojection(prob, oracle)
    prob["poses"] = perturb_poses(prob, sigma_t=0.005, sigma_r=0.002)
    return prob, tgt


def update_graph():
    synth.CONFIGS["tiny"] = (6, 4, 6, 3)
    ii, jj, kk = synth.replay_graph(6, 4, 6, 7)
    return ii, jj, kk


def update_inputs(device="cpu"):
    ii, jj, kk = update_graph()
    E = len(ii)
    g = torch.Generator().manual_seed(77)
    out = dict(net=0.5 * torch.randn(1, E, 384, generator=g), inp=0.5 * torch.randn(1, E, 384, generator=g),
               corr=0.3 * torch.randn(1, E, 882, generator=g), ii=torch.from_numpy(ii),
               jj=torch.from_numpy(jj), kk=torch.from_numpy(kk))
    return {k: v.to(device) for k, v in out.items()}


def encoder_inputs(device="cpu", ht=64, wd=96):
    g = torch.Generator().manual_seed(78)
    frames = []
    for _ in range(2):
        ev = torch.poisson(torch.full((1, 1, 5, ht, wd), 0.3), generator=g)
        ev = ev * (torch.randint(0, 2, ev.shape, generator=g) * 2 - 1)
        im = torch.rand(1, 1, 3, ht, wd, generator=g) * 2 - 0.5
        frames.append((ev.to(device), im.to(device)))
    return frames


def selection_events(device="cpu"):
    g = torch.Generator().manual_seed(79)
    ev = torch.poisson(torch.full((1, 1, 5, 480, 640), 0.12), generator=g)
    ys, xs = torch.meshgrid(torch.arange(480.0), torch.arange(640.0), indexing="ij")
    blob = torch.exp(-((xs - 300) ** 2 + (ys - 200) ** 2) / (2 * 150.0 ** 2))
    return (ev * (0.25 + blob)).to(device)


def event_stream(n=120000, ht=96, wd=128, seed=80):
    """raw events in arrival order like data/events.py: x, y uint16, p int8 in {-1, +1}; a hot cluster makes some
    cells collect > 127 events so that the int8 cast (utils/transformers.py:159) wraps"""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, wd, n).astype(np.uint16)
    y = rng.integers(0, ht, n).astype(np.uint16)
    p = (rng.integers(0, 2, n) * 2 - 1).astype(np.int8)
    hot = rng.random(n) < 0.02
    x[hot], y[hot], p[hot] = 7, 5, 1
    return x, y, p, ht, wd


def single_scale_inputs(device="cpu", ht=32, wd=48):
    """three (event stack, image) pairs; the second image is all zero so that the data-dependent
    `image_is_present` branch of ramp/extractor.py:254-258 is exercised"""
    g = torch.Generator().manual_seed(81)
    frames = []
    for f in range(3):
        ev = torch.poisson(torch.full((1, 1, 5, ht, wd), 0.3), generator=g)
        ev = ev * (torch.randint(0, 2, ev.shape, generator=g) * 2 - 1)
        im = torch.rand(1, 1, 3, ht, wd, generator=g) * 2 - 0.5
        if f == 1:
            im = torch.zeros_like(im)
        frames.append((ev.to(device), im.to(device)))
    return frames


def encoder_clip_inputs(device="cpu", T=4, ht=32, wd=48):
    """one CALL with T frames: the per-pixel LSTMs then see sequences of length T (training clips)"""
    g = torch.Generator().manual_seed(82)
    ev = torch.poisson(torch.full((1, T, 5, ht, wd), 0.3), generator=g)
    ev = ev * (torch.randint(0, 2, ev.shape, generator=g) * 2 - 1)
    im = torch.rand(1, T, 3, ht, wd, generator=g) * 2 - 0.5
    return ev.to(device), im.to(device), torch.ones(1, T, dtype=torch.bool)


def pose_pred_graph():
    """toy patch graph with a virtual frame (the input of the pose-prediction fixture, tests/golden/pose_pred.npz)"""
    import torch
    M, nfr = 3, 9
    ii, jj, kk = [], [], []
    for i in range(nfr - 1):
        for p in range(M):
            for j in range(max(0, i - 3), min(nfr - 1, i + 4)):
                ii.append(i); jj.append(j); kk.append(i * M + p)
    for i in range(nfr - 1 - 3, nfr - 1):            # edges to the virtual frame nfr - 1, appended last
        for p in range(M):
            ii.append(i); jj.append(nfr - 1); kk.append(i * M + p)
    g = torch.Generator().manual_seed(7)
    E = len(ii)
    coords = torch.rand(1, E, 2, 3, 3, generator=g) * 200 - 20        # some observations leave the 160 x 120 map
    return {"ii": torch.tensor(ii), "jj": torch.tensor(jj), "kk": torch.tensor(kk), "coords": coords,
            "tstamps": torch.arange(0, 5 * nfr, 5), "next_frame_index": nfr - 1, "data_shape": (120, 160), "M": M}
