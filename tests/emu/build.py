"""Build the host emulator (test infrastructure) with g++.  See pdq_emu.cpp."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libpdq_emu.so")
SRC = os.path.join(HERE, "pdq_emu.cpp")
DEPS = [SRC] + [os.path.join(HERE, "..", "..", "pydeseq2_b200", "csrc", f)
                for f in ("pdq_gene.cuh", "pdq_math.cuh", "pdq_host_linalg.h", "pdq_trend.cuh", "pdq_fast.cuh", "pdq_shrink.cuh")]


def build(force=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    # -ffp-contract=off: fma() calls stay explicit, nothing else is fused, so results do not depend on g++'s mood
    cmd = ["g++", "-O2", "-std=c++20", "-pthread", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", SRC, "-o", OUT, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
