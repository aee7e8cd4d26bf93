"""The reference's CPU-only per-frame path at FULL graph size (TEST / BASELINE INFRASTRUCTURE).

BASELINE.json asks for "the reference's own CPU-only path (PyTorch conv + Python BA, CUDA ops bypassed)" timed
on the host cores.  The reference cannot run on a CPU as shipped (device="cuda" is hard-coded in
ramp/net.py:175-199 and ramp/Ramp_vo.py:24-94; altcorr / fastba are CUDA-only), so one frame is assembled from
the reference's OWN modules wherever they are device-agnostic, and from the numpy/torch restatement (port) only
where the reference has nothing but a CUDA kernel:

  stage        code that runs                                                          kind
  encoder      ramp.extractor.MultiScaleMergerDoubleNet (staged copy, unmodified)      reference
  reproject    ramp.projective_ops.transform on ramp.lietorch.SE3                      reference (+ lietorch shim)
  corr         oracle.ref_vo.corr_pyramid_torch (correlation_kernel.cu:83-136,221-232) port
  update op    ramp.net.Update (unmodified) with torch_scatter shim; fastba.neighbors  reference (+ neighbors port:
               replaced by oracle.ref_ops.neighbors (ba.cpp:59-97 is CPU code but        ba.cpp needs a CUDA tensor)
               returns CUDA tensors)
  BA           ramp.ba.BA — the reference's Python BA, 2 iterations (Ramp_vo.py:304)   reference

Every step processes ALL E edges of the steady-state default.yaml graph: nothing is extrapolated.
"""
import os
import time

import numpy as np
import torch

from . import ref_ops as O
from . import ref_vo


class CpuReferencePath:
    def __init__(self, state_dict, config="default", n_frames=40, seed=0, threads=None):
        from rampvo_b200 import synth        # seeded synthetic inputs only (numpy), no kernels
        from . import ref_gpu_vo as R
        self.cores = threads or os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.kind = "reference" if R.available() else "port"
        self.prob = prob = synth.make_problem(config, n_frames, seed=seed)
        self.E, self.M = prob["E"], prob["M"]
        gmap, pyr = synth.make_features(32, self.M * 32, seed=seed, dtype=np.float32)
        self.gmap = torch.from_numpy(gmap.transpose(0, 3, 1, 2).copy())
        self.pyr = [torch.from_numpy(p.transpose(0, 3, 1, 2).copy()) for p in pyr]
        g = torch.Generator().manual_seed(5)
        self.imap = torch.randn(self.M * 32, 384, generator=g) * 0.1
        self.net = torch.zeros(1, self.E, 384)
        self.seq = synth.SyntheticSequence(seed=seed, device="cpu")
        self.t = {k: torch.from_numpy(prob[k]) for k in ("ii", "jj", "kk")}
        self.poses = torch.from_numpy(prob["poses"])[None]
        self.patches = torch.from_numpy(prob["patches"])[None]
        self.intr = torch.from_numpy(prob["intrinsics"])[None]
        sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
        if self.kind == "reference":
            self.ns = ns = R.load(extensions=True, with_vo=False)
            enc = ns.extractor.MultiScaleMergerDoubleNet(evs_ch_dim=5, img_ch_dim=3, lstm_dim=16, output_dim_f=128,
                                                         output_dim_i=384, norm_fn_fmap="instance",
                                                         norm_fn_imap="none", norm_superstate=False)
            enc.load_state_dict({k[len("patchify.encoder."):]: v for k, v in sd.items()
                                 if k.startswith("patchify.encoder.")}, strict=True)
            self.enc = enc.eval()
            up = ns.net.Update(3)
            up.load_state_dict({k[len("update."):]: v for k, v in sd.items() if k.startswith("update.")}, strict=True)
            self.update = up.eval()
        else:
            self.enc = ref_vo.Encoder(sd)
            self.p_up = {k[len("update."):]: v for k, v in sd.items() if k.startswith("update.")}

    @staticmethod
    def _neighbors(kk, jj):
        a, b = O.neighbors(kk.numpy(), jj.numpy())
        return torch.from_numpy(a), torch.from_numpy(b)

    @torch.no_grad()
    def step(self, s):
        """one frame: encoder on a fresh 640x480 input + one recurrent update over all E edges; returns the
        wall-time split in seconds"""
        tm = {}
        ev, im, mask = self.seq.frame(s)
        t = time.perf_counter()
        if self.kind == "reference":
            self.enc(events=ev, images=im, mask=torch.tensor([True]), reinit_hidden=(s == 0))
        else:
            self.enc(ev, im, [True], reinit_hidden=(s == 0))
        tm["encoder"] = time.perf_counter() - t
        ii, jj, kk = self.t["ii"], self.t["jj"], self.t["kk"]
        M = self.M
        t = time.perf_counter()
        if self.kind == "reference":
            SE3 = self.ns.lietorch.SE3
            coords = self.ns.pops.transform(SE3(self.poses), self.patches, self.intr, ii, jj, kk)
            coords = coords.permute(0, 1, 4, 2, 3).contiguous()            # Ramp_vo.py:192
        else:
            c = O.transform(self.prob["poses"], self.prob["patches"], self.prob["intrinsics"], self.prob["ii"],
                            self.prob["jj"], self.prob["kk"], dtype=np.float32)[0]
            coords = torch.from_numpy(np.ascontiguousarray(c.transpose(0, 3, 1, 2))).float()[None]
        tm["reproject"] = time.perf_counter() - t
        t = time.perf_counter()
        corr = ref_vo.corr_pyramid_torch(self.gmap, self.pyr, coords[0], kk % (M * 32), jj % 32)[None]
        tm["corr"] = time.perf_counter() - t
        t = time.perf_counter()
        ctx = self.imap[kk % (M * 32)][None]
        if self.kind == "reference":
            real = self.ns.net.fastba
            self.ns.net.fastba = type("fastba_cpu", (), {"neighbors": staticmethod(self._neighbors)})
            try:
                net, (delta, weight, _) = self.update(self.net, ctx, corr, None, ii, jj, kk)
            finally:
                self.ns.net.fastba = real
        else:
            n2, d, w = ref_vo.update_operator(self.p_up, self.net[0], ctx[0], corr[0], ii, jj, kk)
            net, delta, weight = n2[None], d[None], w[None]
        self.net = net
        tm["update_op"] = time.perf_counter() - t
        t = time.perf_counter()
        target = coords[..., 1, 1] + delta.float()
        if self.kind == "reference":
            SE3 = self.ns.lietorch.SE3
            poses, patches = SE3(self.poses.clone()), self.patches.clone()
            lm = 1e-4                                                       # a float broadcasts (ba.py:155-158)
            for _ in range(2):                                              # iterations=2 (Ramp_vo.py:304)
                poses, patches = self.ns.ba.BA(poses, patches, self.intr, target, weight.float(), lm, ii, jj, kk,
                                               bounds=[-64, -64, 160 + 64, 120 + 64], ep=1.0, fixedp=self.prob["t0"])
        else:
            ref_vo.python_ba(self.prob["poses"], self.prob["patches"], self.prob["intrinsics"], target[0].numpy(),
                             weight[0].numpy(), 1e-4, self.prob["ii"], self.prob["jj"], self.prob["kk"],
                             self.prob["t0"], self.prob["t1"], 2)
        tm["ba"] = time.perf_counter() - t
        return tm


def run(state_dict, steps, warmup, budget_s=150.0, config="default"):
    """-> dict(per_frame_s, fps, steps_done, stages, kind, cores, E).  Warm-up and timed steps are real full-E
    frames; the run stops early when the wall-clock budget is used up (at least one timed step)."""
    path = CpuReferencePath(state_dict, config)
    t_all = time.perf_counter()
    times, stages = [], []
    for s in range(warmup):
        path.step(s)
        if time.perf_counter() - t_all > 0.3 * budget_s:
            warmup = s + 1
            break
    for s in range(warmup, warmup + max(1, steps)):
        t0 = time.perf_counter()
        tm = path.step(s)
        times.append(time.perf_counter() - t0)
        stages.append(tm)
        if time.perf_counter() - t_all > budget_s:
            break
    per_frame = float(np.mean(times))
    st = {k: float(np.mean([x[k] for x in stages])) for k in stages[0]}
    return {"per_frame_s": per_frame, "fps": 1.0 / per_frame, "steps_done": len(times), "warmup_done": warmup,
            "stages_s": st, "kind": path.kind, "cores": path.cores, "E": path.E}
