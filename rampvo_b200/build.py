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
    cmd = [nvcc, "-arch=sm_100a", "-shared", "-o", LIB] + objs     # no default-arch (sm_52) link stub
    subprocess.check_call(cmd)
    with open(STAMP, "w") as fh:
        fh.write(fp)
    return LIB


def sass_summary(out_path=None):
    """Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses (tcgen05 MMA = UTCHMMA /
    UTCQMMA, TMEM loads = LDTM, TMA = UTMALDG / UBLKCP, Ampere-style tensor cores = HMMA, cp.async = LDGSTS), from
    `cuobjdump -sass` of the built library.  Written to profiles/sass_summary.txt by __graft_entry__.build()."""
    import re
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    txt = subprocess.run([cuobjdump, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    keys = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDGSTS", "SYNCS")
    rows, cur, cnt, arch = [], None, None, set()
    filt = shutil.which("c++filt")
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur:
                rows.append((cur, cnt))
            cur, cnt = m.group(1), dict.fromkeys(keys, 0)
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        if cur:
            for k in keys:
                if re.search(r"\b%s(\.|\b)" % k, line):
                    cnt[k] += 1
    if cur:
        rows.append((cur, cnt))
    names = [r[0] for r in rows]
    if filt and names:
        dem = subprocess.run([filt], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        if len(dem) == len(names):
            names = [d.split("(")[0][-60:] for d in dem]
    lines = ["# SASS summary of rampvo_b200/librampvo_b200.so (cuobjdump -sass; code objects: %s)" % ", ".join(sorted(arch)),
             "# %-58s %s" % ("kernel", " ".join("%8s" % k for k in keys))]
    for n, (_, c) in sorted(zip(names, rows), key=lambda x: x[0]):
        if any(c.values()):
            lines.append("%-60s %s" % (n, " ".join("%8d" % c[k] for k in keys)))
    out = "\n".join(lines) + "\n"
    if out_path:
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        with open(out_path, "w") as fh:
            fh.write(out)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
