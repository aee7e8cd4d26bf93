"""Generates tests/golden/pose_pred.npz by running the reference's own host functions
(ramp/pose_prediction/pose_pred_utils.py: compute_patch_track__, fit_model_patch_track — the CPU-capable half of the
pose-prediction branch) on a seeded toy graph.  Run in the build container: python tests/golden/make_pose_pred_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden_inputs import pose_pred_graph  # noqa: E402

REF = os.environ.get("RVO_REFERENCE", "/root/reference")


def main():
    src = open(os.path.join(REF, "ramp", "pose_prediction", "pose_pred_utils.py")).read()
    ns = {}
    pre = ("from collections import defaultdict\nimport numpy as np\nimport torch\n"
           "from scipy.interpolate import UnivariateSpline\n")
    # the module imports sklearn / matplotlib at the top; only the two pure functions are executed
    exec(pre + src[src.index("def compute_patch_track__"):src.index("def motion_bootstrap")], ns)
    exec(src[src.index("def fit_model_patch_track"):src.index("def predict_patch_on_model")], ns)
    g = pose_pred_graph()
    tracks = ns["compute_patch_track__"](g["coords"], g["ii"], g["jj"], g["kk"], g["next_frame_index"])
    models = ns["fit_model_patch_track"](g["next_frame_index"], tracks, g["tstamps"], g["ii"], g["jj"], g["data_shape"],
                                         frequency=30, deg=3)
    keys = [k for k in tracks if len(tracks[k]) > 0]
    out = {"keys": np.asarray(keys, dtype=np.int64)}
    for n, k in enumerate(keys):
        out["track_%d" % n] = tracks[k].numpy()
        sx, sy, w, last_t = models[k]
        out["pred_%d" % n] = np.asarray([[float(sx(last_t + s / 30.0)), float(sy(last_t + s / 30.0))] for s in (1, 2, 5)])
        out["w_%d" % n] = np.asarray([w, last_t], dtype=np.float64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pose_pred.npz"), **out)
    print("wrote pose_pred.npz with", len(keys), "tracks")


if __name__ == "__main__":
    main()
