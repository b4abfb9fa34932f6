"""Build the CPU oracle (TEST INFRASTRUCTURE ONLY) into oracle/_build/liborp_oracle.so.

gcc only; -ffp-contract=off so every fp operation is separately rounded, which is what the
reference's x86-64 builds do and what the CUDA kernels reproduce with --fmad=false.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "liborp_oracle.so")
SOURCES = ["oracle_geom.c", "oracle_minarearect.c", "oracle_dcn.c", "oracle_convex_iou.c"]


def build(force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    deps = srcs + [os.path.join(HERE, "polyclip_body.inc")]
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps)):
        return LIB
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-fopenmp", "-Wall", "-Wno-unused-function", "-o", LIB] + srcs + ["-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
