"""Recipe that compiles the REFERENCE's own CUDA ops into oracle/_ref/ (test infrastructure only).

Nothing from /root/reference is copied into the repository: the sources are read where they lie,
staged under /tmp, minimally patched THERE so that they build against torch 2.11 without Eigen, and
only the two resulting extension modules land in oracle/_ref/ (git-ignored, shipped to the GPU box):

  cuda_corr_ref  <- ramp/altcorr/correlation.cpp + correlation_kernel.cu
      patch: `x.type()` -> `x.scalar_type()` inside the 4 AT_DISPATCH_* macros
      (correlation_kernel.cu:211,273,299,325; torch 2.x removed the DeprecatedTypeProperties overload)
  cuda_ba_ref    <- ramp/fastba/ba_cuda.cu (unmodified) + block_e.cu (unmodified, compiled against
      oracle/eigen_stub/Eigen/Core, a 20-line stand-in for the one Eigen::Array it uses) + ba.cpp
      patch: the Eigen sparse `solve` / `solve_system` (ba.cpp:99-180, never called by the hot path)
      and their includes / pybind entry are dropped; `forward`, `neighbors`, `reproject` are untouched.

The kernels and their host drivers — the algorithms parity is measured against — are compiled
unmodified.  Only tests/ and bench.py's reference timing import these modules.
"""
import os
import re
import shutil
import sys

REF = os.environ.get("RVO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
STAGE = "/tmp/rvo_ref_stage"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _stage():
    shutil.rmtree(STAGE, ignore_errors=True)
    os.makedirs(STAGE)
    a, f = os.path.join(REF, "ramp", "altcorr"), os.path.join(REF, "ramp", "fastba")
    for src in [os.path.join(a, "correlation.cpp"), os.path.join(a, "correlation_kernel.cu"),
                os.path.join(f, "ba.cpp"), os.path.join(f, "ba_cuda.cu"),
                os.path.join(f, "block_e.cu"), os.path.join(f, "block_e.cuh")]:
        shutil.copy(src, STAGE)
    # patch 1: AT_DISPATCH on scalar_type()
    p = os.path.join(STAGE, "correlation_kernel.cu")
    s = open(p).read()
    s, n = re.subn(r"(AT_DISPATCH_FLOATING_TYPES_AND_HALF\(\s*\w+)\.type\(\)", r"\1.scalar_type()", s)
    assert n == 4, "expected 4 AT_DISPATCH sites, found %d" % n
    open(p, "w").write(s)
    # patch 2: drop the Eigen sparse solver from ba.cpp
    p = os.path.join(STAGE, "ba.cpp")
    s = open(p).read()
    s = s.replace("#include <Eigen/Core>\n", "").replace("#include <Eigen/Sparse>\n", "")
    i0 = s.index("typedef Eigen::SparseMatrix<double> SpMat;")
    i1 = s.index("PYBIND11_MODULE")
    s = s[:i0] + s[i1:]
    s, n = re.subn(r'\n\s*m\.def\("solve_system"[^\n]*\n', "\n", s)
    assert n == 1
    open(p, "w").write(s)


PY_OUT = os.path.join(OUT, "ref_py")


def stage_python():
    """Copies the reference's Python package `ramp/` (UNMODIFIED .py files only) into oracle/_ref/ref_py/ramp so
    that oracle/ref_gpu_vo.py can import the reference's own Ramp_vo / VONet / Update / extractor / projective_ops
    / ba on the GPU box, which has no /root/reference.  oracle/_ref is git-ignored: no reference source enters the
    repository history; the copy ships to the box like the compiled .so files."""
    if not os.path.isdir(REF):
        raise RuntimeError("reference checkout not found at %s" % REF)
    dst = os.path.join(PY_OUT, "ramp")
    shutil.rmtree(PY_OUT, ignore_errors=True)
    src = os.path.join(REF, "ramp")
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ("src", "include", "__pycache__")]
        for f in files:
            if f.endswith(".py"):
                rel = os.path.relpath(os.path.join(root, f), src)
                os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
                shutil.copy(os.path.join(root, f), os.path.join(dst, rel))
                n += 1
    # the 5-bin event stack builder (utils/transformers.py:128-161) is the checker of the GPU event-stacking kernel
    os.makedirs(os.path.join(PY_OUT, "ref_utils"), exist_ok=True)
    shutil.copy(os.path.join(REF, "utils", "transformers.py"), os.path.join(PY_OUT, "ref_utils", "transformers.py"))
    return n


def build(verbose=False):
    if not os.path.isdir(REF):
        raise RuntimeError("reference checkout not found at %s" % REF)
    stage_python()
    import torch  # noqa: F401
    from torch.utils.cpp_extension import load
    _stage()
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    specs = [
        ("cuda_corr_ref", ["correlation.cpp", "correlation_kernel.cu"], []),
        ("cuda_ba_ref", ["ba.cpp", "ba_cuda.cu", "block_e.cu"], [os.path.join(HERE, "eigen_stub")]),
    ]

    def one(spec):
        name, srcs, inc = spec
        bdir = os.path.join("/tmp", "rvo_ref_build_" + name)
        os.makedirs(bdir, exist_ok=True)
        mod = load(name=name, sources=[os.path.join(STAGE, x) for x in srcs],
                   extra_include_paths=inc, extra_cflags=["-O3"],
                   extra_cuda_cflags=["-O3"] + ARCH, build_directory=bdir, verbose=verbose)
        shutil.copy(os.path.join(bdir, name + ".so"), os.path.join(OUT, name + ".so"))
        return name, mod

    # the two extensions are independent (~5 minutes of nvcc each): build them side by side
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=2) as ex:
        mods = dict(ex.map(one, specs))
    return mods


def load_ref(name):
    """Import a prebuilt module from oracle/_ref (returns None when it was never built)."""
    import importlib.util
    import torch  # noqa: F401  (must be imported before a torch extension)
    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    if "--python-only" in sys.argv:
        print(stage_python(), "python files staged")
        sys.exit(0)
    build(verbose="-v" in sys.argv)
    print(sorted(os.listdir(OUT)))
