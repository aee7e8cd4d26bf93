"""Builds librampvo_b200.so (the C-ABI library, include/rampvo_b200.h) in-tree with nvcc for sm_100a.

The library is plain CUDA C++ (no torch headers) so that the C-ABI boundary stays a C boundary;
`python -m rampvo_b200.build` or `__graft_entry__.build()` runs this.  nvcc cross-compiles without
a GPU.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librampvo_b200.so")
STAMP = os.path.join(HERE, ".librampvo_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-ftemplate-depth=4096",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    root = os.path.dirname(HERE)
    files = _sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    files.append(os.path.join(root, "include", "rampvo_b200.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: librampvo_b200.so cannot be built")
    return nvcc


def build_library(force=False, verbose=False, debug=False):
    """Compile every .cu under csrc/ into one shared library; returns its path.  debug=True adds -DRVO_DEBUG
    (the RVO_CORR_DBG / RVO_UP_TRACE isolation switches of tools/corr_bench.py and tools/gemm_bench.py; with them
    set the kernels produce WRONG results by design, so the release library compiles them out)."""
    fp = _fingerprint() + ("-debug" if debug else "")
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == fp:
                return LIB
    nvcc = find_nvcc()
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    for src in _sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        cmd = ([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DRVO_DEBUG"] if debug else []) +
               ["-c", src, "-o", obj])
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode(errors="replace"))
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [nvcc, "-shared", "-o", LIB] + objs
    subprocess.check_call(cmd)
    with open(STAMP, "w") as fh:
        fh.write(fp)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
