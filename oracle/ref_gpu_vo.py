"""The REFERENCE itself on the GPU (TEST / BASELINE INFRASTRUCTURE — never imported by rampvo_b200/).

`load()` imports the reference's own, unmodified Python package `ramp` (staged by oracle/build_ref.py into
oracle/_ref/ref_py/, git-ignored, shipped to the GPU box like the .so files) on top of

  cuda_corr, cuda_ba      the reference's CUDA ops compiled from /root/reference (oracle/_ref/*.so; the
                          recipe and its three documented patches are in oracle/build_ref.py)
  lietorch_backends       oracle/shims/lietorch_backends.py (Eigen is absent -> the extension cannot be built;
                          forward SE3/SO3 entry points restated as tensor arithmetic)
  torch_scatter           oracle/shims/torch_scatter.py (absent from the image, unpinned upstream)
  evo, matplotlib, h5py, hdf5plugin, yacs, data
                          empty stand-ins: imported at module scope by ramp/utils.py:1-18 and
                          ramp/pose_prediction/*.py, never touched by the tracking path

so that ramp.Ramp_vo.Ramp_vo (ramp/Ramp_vo.py:27-410), ramp.net.VONet, ramp.altcorr, ramp.fastba,
ramp.projective_ops run exactly as upstream runs them: this is (i) the END-TO-END PARITY ORACLE of
tests/test_gpu_e2e_reference.py and (ii) the same-box reference-GPU timing bench.py reports as `ref_gpu`
(the "reference single-GPU evaluate.py frames/sec" of BASELINE.json's north star; evaluate.run's loop,
evaluate.py:247-255, is `run_sequence` below).
"""
import importlib
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
PY_DIR = os.path.join(REF_DIR, "ref_py")

_loaded = None


def available():
    return (os.path.isdir(os.path.join(PY_DIR, "ramp")) and os.path.exists(os.path.join(REF_DIR, "cuda_corr_ref.so"))
            and os.path.exists(os.path.join(REF_DIR, "cuda_ba_ref.so")))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []            # lets `import name.sub` resolve to further stubs
    sys.modules[name] = m
    return m


def _load_ext(name, alias):
    import torch  # noqa: F401
    path = os.path.join(REF_DIR, name + ".so")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[alias] = mod
    return mod


def install(extensions=True):
    """Registers the stand-ins and (optionally) the compiled reference ops under the names the reference imports."""
    from .shims import lietorch_backends, torch_scatter
    sys.modules["lietorch_backends"] = lietorch_backends
    sys.modules["torch_scatter"] = torch_scatter
    if extensions:
        _load_ext("cuda_corr_ref", "cuda_corr")      # ramp/altcorr/correlation.py:2
        _load_ext("cuda_ba_ref", "cuda_ba")          # ramp/fastba/ba.py:2
    for name in ("evo", "evo.core", "h5py", "hdf5plugin", "matplotlib", "yacs"):
        if name not in sys.modules:
            _stub(name)
    _stub("evo.core.trajectory", PoseTrajectory3D=object)
    _stub("matplotlib.pyplot")
    _stub("data", H5EventHandle=object)              # ramp/utils.py:18 `from data import H5EventHandle`

    class _CN(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    _stub("yacs.config", CfgNode=_CN)
    if PY_DIR not in sys.path:
        sys.path.insert(0, PY_DIR)


def load(extensions=True, with_vo=True):
    """-> namespace with the reference modules (ramp.net, ramp.Ramp_vo, ...).  with_vo needs a GPU:
    ramp/Ramp_vo.py:24 builds a CUDA tensor at import time."""
    global _loaded
    if _loaded is not None and (not with_vo or hasattr(_loaded, "Ramp_vo")):
        return _loaded
    if not os.path.isdir(os.path.join(PY_DIR, "ramp")):
        raise RuntimeError("oracle/_ref/ref_py is missing: run `python oracle/build_ref.py` where /root/reference exists")
    install(extensions)
    ns = types.SimpleNamespace()
    ns.net = importlib.import_module("ramp.net")
    ns.extractor = importlib.import_module("ramp.extractor")
    ns.blocks = importlib.import_module("ramp.blocks")
    ns.pops = importlib.import_module("ramp.projective_ops")
    ns.ba = importlib.import_module("ramp.ba")
    ns.lietorch = importlib.import_module("ramp.lietorch")
    ns.utils = importlib.import_module("ramp.utils")
    if extensions:
        ns.altcorr = importlib.import_module("ramp.altcorr")
        ns.fastba = importlib.import_module("ramp.fastba")
    if with_vo:
        ns.Ramp_vo = importlib.import_module("ramp.Ramp_vo")
    _loaded = ns
    return ns


def make_vo(cfg, state_dict, train_cfg, ht=480, wd=640):
    """reference Ramp_vo with the given weights (state-dict keys are shared with rampvo_b200.net.VONet)"""
    ns = load()
    net = ns.net.VONet(train_cfg)
    net.load_state_dict(state_dict, strict=True)
    return ns.Ramp_vo.Ramp_vo(cfg, net, train_cfg, ht=ht, wd=wd)


def run_sequence(vo, frames, intrinsics, final_updates=0):
    """evaluate.run (evaluate.py:247-255): one slam() call per (events, image, mask), then extra updates"""
    for t, (ev, im, mask) in enumerate(frames):
        vo(t, input_tensor=(ev, im, mask), intrinsics=intrinsics)
    for _ in range(final_updates):
        vo.update()
    return vo
