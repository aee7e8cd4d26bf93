"""fastba on the precise.yaml graph (E = 660 600, 30 free poses: BASELINE.json configs[2]) and on default.yaml:
CUDA-event time of fastba.BA for iterations in {1, 2, 4} (eff_impl False / True give the same kernels) next to the
reference's cuda_ba when oracle/_ref is present.  Under `ncu --metrics gpu__time_duration.sum` this is the fastba
kernel sweep.  Usage: python tools/ba_bench.py [precise|default]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_ops as O  # noqa: E402  (input generation / baseline only)
from rampvo_b200 import fastba, synth  # noqa: E402
from tests.util import perturb_poses, problem_tensors, targets_from_reprojection  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "precise"
    prob = synth.make_problem(cfg, {"default": 40, "precise": 80}[cfg], seed=5)
    tgt = targets_from_reprojection(prob, O)
    prob["poses"] = perturb_poses(prob)
    tg = torch.from_numpy(tgt).cuda()[None]
    wg = torch.from_numpy(prob["weight"]).cuda()[None]
    lm = torch.tensor([1e-4], device="cuda")
    ref = build_ref.load_ref("cuda_ba_ref")
    rows = []
    for iters in (1, 2, 4):
        for eff in (False, True):
            def ours():
                b = problem_tensors(prob)
                torch.cuda.synchronize()
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fastba.BA(b["poses"], b["patches"], b["intrinsics"], tg, wg, lm, b["ii"], b["jj"], b["kk"], prob["t0"],
                          prob["t1"], prob["M"], iters, eff_impl=eff)
                c.record()
                torch.cuda.synchronize()
                return a.elapsed_time(c) * 1e3

            def theirs():
                b = problem_tensors(prob)
                torch.cuda.synchronize()
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ref.forward(b["poses"], b["patches"], b["intrinsics"], tg, wg, lm, b["ii"], b["jj"], b["kk"], prob["M"],
                            prob["t0"], prob["t1"], iters, eff)
                c.record()
                torch.cuda.synchronize()
                return a.elapsed_time(c) * 1e3
            ours()
            t = sorted(ours() for _ in range(5))[2]
            r = None
            if ref is not None:
                theirs()
                r = sorted(theirs() for _ in range(3))[1]
            rows.append({"config": cfg, "E": prob["E"], "free_poses": prob["t1"] - prob["t0"], "iterations": iters,
                         "eff_impl": eff, "ours_us": round(t, 1), "cuda_ba_ref_us": None if r is None else round(r, 1)})
            print(json.dumps(rows[-1]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ba_bench_%s.json" % cfg), "w") as fh:
        json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
