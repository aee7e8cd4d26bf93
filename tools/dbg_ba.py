import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import ref_ops as O, build_ref
from rampvo_b200 import fastba, synth
from tests.util import *
ref_ba = build_ref.load_ref("cuda_ba_ref")
for cfg, nf, t0 in [("cfg1", 8, None), ("cfg1", 8, 4), ("default", 40, None)]:
    prob = synth.make_problem(cfg, nf, seed=3)
    if t0 is not None: prob["t0"] = t0
    tgt = targets_from_reprojection(prob, O)
    prob["poses"] = perturb_poses(prob)
    tg = torch.from_numpy(tgt).cuda()[None]; wg = torch.from_numpy(prob["weight"]).cuda()[None]
    lm = torch.tensor([1e-4], device="cuda")
    for iters in (1, 2):
        a = problem_tensors(prob); b = problem_tensors(prob)
        ref_ba.forward(a["poses"], a["patches"], a["intrinsics"], tg, wg, lm, a["ii"], a["jj"], a["kk"], prob["M"], prob["t0"], prob["t1"], iters, False)
        fastba.BA(b["poses"], b["patches"], b["intrinsics"], tg, wg, lm, b["ii"], b["jj"], b["kk"], prob["t0"], prob["t1"], prob["M"], iters)
        pe, qe = O.ba(prob["poses"], prob["patches"], prob["intrinsics"], tgt, prob["weight"], 1e-4, prob["ii"], prob["jj"], prob["kk"], prob["t0"], prob["t1"], iterations=iters)
        p32, q32 = O.ba(prob["poses"], prob["patches"], prob["intrinsics"], tgt, prob["weight"], 1e-4, prob["ii"], prob["jj"], prob["kk"], prob["t0"], prob["t1"], iterations=iters, dtype=np.float32)
        A = a["poses"][0].cpu().numpy(); B = b["poses"][0].cpu().numpy()
        s = O.ba_system(prob["poses"], prob["patches"], prob["intrinsics"], tgt, prob["weight"], 1e-4, prob["ii"], prob["jj"], prob["kk"], prob["t0"], prob["t1"])
        S = s["S"].copy(); S[np.diag_indices_from(S)] += 1e-4*np.diag(S)+1
        print(cfg, prob["t0"], iters, "mine-vs-ref %.2e mine-vs-f64 %.2e ref-vs-f64 %.2e f32oracle-vs-f64 %.2e cond %.2e" % (rel_err(B, A), rel_err(B, pe), rel_err(A, pe), rel_err(p32, pe), np.linalg.cond(S)),
              "depth: mine-vs-ref %.2e mine-vs-f64 %.2e ref-vs-f64 %.2e" % (rel_err(b["patches"][0,:,2].cpu().numpy(), a["patches"][0,:,2].cpu().numpy()), rel_err(b["patches"][0,:,2].cpu().numpy(), qe[:,2]), rel_err(a["patches"][0,:,2].cpu().numpy(), qe[:,2])))
