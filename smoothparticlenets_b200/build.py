"""Builds libspnb.so -- the sm_100a CUDA library behind the C ABI in include/spnb.h.

    python -m smoothparticlenets_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library is built IN-TREE (next to this file) so that it
travels with the repository snapshot to the GPU box; it is git-ignored.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libspnb.so")
SOURCES = ["common.cu", "hashgrid.cu", "convsp.cu", "convsp_small.cu", "convsp_group.cu", "convsdf.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # Separately rounded IEEE float ops in source order: geometric predicates must agree with the
    # reference's CPU build bit for bit (no FMA contraction, precise div/sqrt, no fast-math).
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-shared",
]


def _nvcc():
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "spnb.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, extra_flags=()):
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags)
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    extra = ["-Xptxas", "-v"] if "--ptxas" in sys.argv else []
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv or bool(extra),
                        extra_flags=extra))
