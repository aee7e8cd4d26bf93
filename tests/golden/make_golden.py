"""Generates the golden fixtures under tests/golden/ by RUNNING THE REFERENCE'S OWN PYTHON CODE from
/root/reference (authoring container only; the GPU box has no reference checkout, so the outputs are
committed as small .npz files next to this script).

    python tests/golden/make_golden.py

How each reference piece is executed unmodified on a CPU-only box:
  ramp/extractor.py, ramp/blocks.py, ramp/projective_ops.py, ramp/ba.py
      loaded from their files as members of a synthetic package `ramp` whose missing siblings are
      stand-ins: `torch_scatter` (scatter_sum / scatter_softmax written with index_add / scatter_reduce;
      the real package is unpinned in requirements.txt:4 and absent here), `ramp.lietorch.SE3` (this
      repo's tensor SE3 layer, itself checked against the reference's lietorch identities), `ramp.fastba`
      with `neighbors` from the numpy oracle (the real one is a CUDA extension), `ramp.utils.Timer`.
  ramp/net.py: class Update, ramp/utils.py: nms_image + get_coords_from_topk_events
      the class / function source is cut out of the file with `ast` and exec'd as is; the only edit
      is device="cuda" -> "cpu" in get_coords_from_topk_events (utils.py:216).
Inputs are regenerated from seeds by the tests (rampvo_b200.synth and torch.manual_seed), weights are
the seeded initialisation of THIS repo's modules loaded into the reference modules with
strict=True — which also proves the state-dict key compatibility.
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("RVO_REFERENCE", "/root/reference")

from oracle import ref_ops as O  # noqa: E402
from rampvo_b200 import lietorch as my_lietorch  # noqa: E402
from rampvo_b200 import synth  # noqa: E402
from tests import golden_inputs as GI  # noqa: E402


def _install_shims():
    ts = types.ModuleType("torch_scatter")

    def scatter_sum(src, index, dim=1, dim_size=None):
        assert dim == 1
        n = int(index.max()) + 1 if dim_size is None else dim_size
        out = torch.zeros(src.shape[0], n, *src.shape[2:], dtype=src.dtype)
        return out.index_add_(1, index, src)

    def scatter_softmax(src, index, dim=1):
        assert dim == 1
        n = int(index.max()) + 1
        idx = index.view(1, -1, *([1] * (src.dim() - 2))).expand_as(src)
        mx = torch.full((src.shape[0], n) + tuple(src.shape[2:]), -float("inf"), dtype=src.dtype)
        mx = mx.scatter_reduce(1, idx, src, reduce="amax", include_self=True)
        ex = (src - mx[:, index]).exp()
        den = torch.zeros_like(mx).index_add_(1, index, ex)
        return ex / den[:, index]
    ts.scatter_sum, ts.scatter_softmax = scatter_sum, scatter_softmax
    sys.modules["torch_scatter"] = ts

    pkg = types.ModuleType("ramp")
    pkg.__path__ = [os.path.join(REF, "ramp")]
    sys.modules["ramp"] = pkg
    lt = types.ModuleType("ramp.lietorch")
    lt.SE3 = my_lietorch.SE3
    sys.modules["ramp.lietorch"] = lt
    pkg.lietorch = lt
    fb = types.ModuleType("ramp.fastba")

    def neighbors(ii, jj):
        a, b = O.neighbors(ii.numpy(), jj.numpy())
        return torch.from_numpy(a), torch.from_numpy(b)
    fb.neighbors = neighbors
    sys.modules["ramp.fastba"] = fb
    pkg.fastba = fb
    ut = types.ModuleType("ramp.utils")

    class Timer:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False
    ut.Timer = Timer
    sys.modules["ramp.utils"] = ut
    pkg.utils = ut


def _load(name):
    spec = importlib.util.spec_from_file_location("ramp." + name, os.path.join(REF, "ramp", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ramp." + name] = mod
    spec.loader.exec_module(mod)
    setattr(sys.modules["ramp"], name, mod)
    return mod


def _cut(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            out[node.name] = ast.get_source_segment(src, node)
    return out


def main():
    torch.set_num_threads(8)
    _install_shims()
    pops = _load("projective_ops")
    blocks = _load("blocks")
    extractor = _load("extractor")
    refba = _load("ba")

    # ---- (a) projective_ops.transform with Jacobians, float64
    prob = GI.pops_problem()
    t = GI.as_torch(prob, torch.float64)
    x1, v, (Ji, Jj, Jz) = pops.transform(my_lietorch.SE3(t["poses"]), t["patches"], t["intrinsics"],
                                         t["ii"], t["jj"], t["kk"], jacobian=True)
    x2 = pops.transform(my_lietorch.SE3(t["poses"]), t["patches"], t["intrinsics"], t["ii"], t["jj"],
                        t["kk"], tonly=True)
    fm = pops.flow_mag(my_lietorch.SE3(t["poses"]), t["patches"], t["intrinsics"], t["ii"], t["jj"],
                       t["kk"], beta=0.5)
    pc = pops.point_cloud(my_lietorch.SE3(t["poses"]), t["patches"], t["intrinsics"],
                          torch.arange(prob["patches"].shape[0]) // prob["M"])
    np.savez_compressed(os.path.join(HERE, "pops_transform.npz"), coords=x1[0].numpy(), valid=v[0].numpy(),
                        Ji=Ji[0].numpy(), Jj=Jj[0].numpy(), Jz=Jz[0].numpy(), coords_tonly=x2[0].numpy(),
                        flow_mag=fm[0].numpy(), points=(pc[0, :, 1, 1, :3] / pc[0, :, 1, 1, 3:]).numpy())

    # ---- (b) the reference's Python BA (ramp/ba.py), ep = 1 to match cuda_ba's damping
    prob, tgt = GI.ba_problem(O)
    t = GI.as_torch(prob, torch.float64)
    poses, patches = my_lietorch.SE3(t["poses"]), t["patches"]
    tg = torch.from_numpy(tgt).double()[None]
    wg = torch.from_numpy(prob["weight"]).double()[None]
    outs = {}
    for it in (1, 2):
        poses, patches = refba.BA(poses, patches, t["intrinsics"], tg, wg, 1e-4, t["ii"], t["jj"], t["kk"],
                                  bounds=[-64, -64, 160 + 64, 120 + 64], ep=1.0, fixedp=prob["t0"])
        outs["poses_%d" % it] = poses.data[0].numpy().copy()
        outs["disps_%d" % it] = patches[0, :, 2, 0, 0].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "python_ba.npz"), **outs)

    # ---- (c) Update operator (ramp/net.py:34-90), fp32 on CPU
    from rampvo_b200.net import Update as MyUpdate
    ns = {"torch": torch, "nn": torch.nn, "GatedResidual": blocks.GatedResidual, "SoftAgg": blocks.SoftAgg,
          "GradientClip": blocks.GradientClip, "fastba": sys.modules["ramp.fastba"], "DIM": 384}
    exec(_cut(os.path.join(REF, "ramp", "net.py"), {"Update"})["Update"], ns)
    torch.manual_seed(GI.UPDATE_SEED)
    mine = MyUpdate(3)
    ref_update = ns["Update"](3)
    ref_update.load_state_dict(mine.state_dict(), strict=True)
    ref_update.eval()
    g = GI.update_inputs()
    with torch.no_grad():
        net, (d, w, _) = ref_update(g["net"], g["inp"], g["corr"], None, g["ii"], g["jj"], g["kk"])
    np.savez_compressed(os.path.join(HERE, "update_op.npz"), net=net[0].numpy(), delta=d[0].numpy(),
                        weight=w[0].numpy())

    # ---- (d) MultiScale encoder (ramp/extractor.py:468-566), two consecutive frames, fp32 CPU
    from rampvo_b200.extractor import MultiScaleMergerDoubleNet as MyEnc
    torch.manual_seed(GI.ENCODER_SEED)
    mine = MyEnc(5, 3)
    ref_enc = extractor.MultiScaleMergerDoubleNet(evs_ch_dim=5, img_ch_dim=3, lstm_dim=16, output_dim_f=128,
                                                  output_dim_i=384, norm_fn_fmap="instance",
                                                  norm_fn_imap="none", norm_superstate=False)
    ref_enc.load_state_dict(mine.state_dict(), strict=True)
    ref_enc.eval()
    outs = {}
    with torch.no_grad():
        for f, (ev, im) in enumerate(GI.encoder_inputs()):
            fmap, imap = ref_enc(events=ev, images=im, mask=torch.tensor([True]), reinit_hidden=(f == 0))
            outs["fmap_%d" % f] = fmap[0, 0].numpy().astype(np.float32)
            outs["imap_%d" % f] = imap[0, 0].numpy().astype(np.float32)
        # an events-only call (mask False) still advances the super state (extractor.py:447-455)
        ev, im = GI.encoder_inputs()[0]
        ref_enc(events=ev, images=im, mask=torch.tensor([False]), reinit_hidden=False)
        ev, im = GI.encoder_inputs()[1]
        fmap, imap = ref_enc(events=ev, images=im, mask=torch.tensor([True]), reinit_hidden=False)
        outs["fmap_after_events_only"] = fmap[0, 0].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "encoder.npz"), **outs)

    # ---- (e) event-biased patch selection (ramp/utils.py:157-226)
    cut = _cut(os.path.join(REF, "ramp", "utils.py"), {"nms_image", "get_coords_from_topk_events"})
    ns = {"torch": torch, "F": torch.nn.functional}
    exec(cut["nms_image"], ns)
    exec(cut["get_coords_from_topk_events"].replace('device="cuda"', 'device="cpu"'), ns)
    ev = GI.selection_events()
    coords = ns["get_coords_from_topk_events"](events=ev, patches_per_image=96, border_suppression_size=0,
                                               non_max_supp_rad=11)
    np.savez_compressed(os.path.join(HERE, "patch_selection.npz"), coords=coords.numpy())
    # ---- (h) MultiScale encoder on a 4-frame CLIP in one call (per-pixel LSTM sequences of length 4), fp32 CPU
    torch.manual_seed(GI.ENCODER_SEED)
    mine = MyEnc(5, 3)
    ref_enc.load_state_dict(mine.state_dict(), strict=True)
    ev, im, mask = GI.encoder_clip_inputs()
    with torch.no_grad():
        fmap, imap = ref_enc(events=ev, images=im, mask=mask, reinit_hidden=True)
    np.savez_compressed(os.path.join(HERE, "encoder_clip.npz"), fmap=fmap[0].numpy().astype(np.float32),
                        imap16=imap[0, :, :16].numpy().astype(np.float32))

    # ---- (g) SingleScale encoder (ramp/extractor.py:187-269), three frames with carried LSTM state, fp32 CPU
    from rampvo_b200.extractor import MergerLSTMsceneEncoder as MySS
    torch.manual_seed(GI.ENCODER_SEED)
    mine = MySS(5, 3)
    ref_ss = extractor.MergerLSTMsceneEncoder(evs_ch_dim=5, img_ch_dim=3, output_lstm_dim=15, output_dim_f=128,
                                              output_dim_i=384, norm_fn_fmap="instance", norm_fn_imap="none",
                                              kernel_size_superstate=1)
    ref_ss.load_state_dict(mine.state_dict(), strict=True)
    ref_ss.eval()
    outs = {}
    with torch.no_grad():
        for f, (ev, im) in enumerate(GI.single_scale_inputs()):
            fmap, imap, _ = ref_ss(events=ev, images=im, reinit_hidden=(f == 0))
            outs["fmap_%d" % f] = fmap[0, 0].numpy().astype(np.float32)
            outs["imap_%d" % f] = imap[0, 0].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "single_scale_encoder.npz"), **outs)

    # ---- (f) 5-bin event stack (utils/transformers.py:128-161 EventToStack_Numpy), run as is
    data_stub = types.ModuleType("data")
    data_stub.Events = object
    sys.modules["data"] = data_stub
    spec = importlib.util.spec_from_file_location("ref_transformers", os.path.join(REF, "utils", "transformers.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)

    class _Ev:
        def __len__(self):
            return len(self.x)
    e = _Ev()
    e.x, e.y, e.p, e.height, e.width = GI.event_stream()
    np.savez_compressed(os.path.join(HERE, "event_stack.npz"), stack=tr.EventToStack_Numpy(5)(e))
    print("golden fixtures written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
