"""Build ``libpydeseq2_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpydeseq2_b200.so")
SOURCES = ["pdq_kernels.cu", "pdq_api.cu"]
HEADERS = ["pdq_math.cuh", "pdq_fast.cuh", "pdq_trend.cuh", "pdq_gene.cuh", "pdq_shrink.cuh", "pdq_internal.h", "pdq_host_linalg.h",
           os.path.join("..", "..", "include", "pydeseq2_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fopenmp", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return nv


def _stale(target: str, deps) -> bool:
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """`defines` / `out` build a tuning variant (e.g. ("-DPDQ_ALPHA_MINB=5",), "libpdq_a5.so") next to the default library."""
    global OUT
    tag = "".join(d.replace("-D", "_").replace("=", "") for d in defines)
    objdir = os.path.join(HERE, "build" + tag)
    target = os.path.join(HERE, out) if out else OUT
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    # the library is newer than every source: nothing to do (the object directory does not travel with gpurun snapshots)
    if not force and not _stale(target, [os.path.join(CSRC, s) for s in SOURCES] + hdrs):
        return target
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [_nvcc()] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode:
            sys.stderr.write(out)
        if pr.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or _stale(target, objs):
        cmd = [_nvcc(), "-shared", "-o", target] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
                                                        "-ldl", "-lrt", "-lpthread", "-lgomp"]
        subprocess.run(cmd, check=True)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
