"""Builds libspnb.so -- the sm_100a CUDA library behind the C ABI in include/spnb.h.

    python -m smoothparticlenets_b200.build [--force] [--verbose] [--ptxas]

nvcc cross-compiles without a GPU.  Sources are compiled to objects in parallel (cached by content +
flags under csrc/_obj/) and linked IN-TREE next to this file so that the library travels with the
repository snapshot to the GPU box; objects and library are git-ignored.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
# SPNB_LIB: build/load a variant library under another name (tuning sweeps prebuilt before a GPU session)
LIB_PATH = os.environ.get("SPNB_LIB") or os.path.join(HERE, "libspnb.so")
SOURCES = ["common.cu", "hashgrid.cu", "convsp.cu", "convsp_small.cu", "convsp_group.cu", "convsp_group_inst_fluid3.cu",
           "convsp_group_inst_fluid2.cu", "convsp_group_inst_single3.cu", "convsp_group_inst_single2.cu",
           "convsp_wide.cu", "convsp_wide_mma.cu", "convsp_wide_bwd.cu", "convsdf.cu", "projection.cu", "fluid_glue.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # Separately rounded IEEE float ops in source order: geometric predicates must agree with the
    # reference's CPU build bit for bit (no FMA contraction, precise div/sqrt, no fast-math).
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
]


def _nvcc():
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def _headers():
    hs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "spnb.h"))
    return hs


def _digest(src, flags):
    h = hashlib.sha1()
    h.update(" ".join(flags).encode())
    for path in [src] + _headers():
        with open(path, "rb") as fp:
            h.update(fp.read())
    return h.hexdigest()[:16]


def _compile(src, flags, verbose):
    name = os.path.splitext(os.path.basename(src))[0]
    obj = os.path.join(OBJ, "%s.%s.o" % (name, _digest(src, flags)))
    if not os.path.exists(obj):
        for old in os.listdir(OBJ):
            if old.startswith(name + "."):
                os.remove(os.path.join(OBJ, old))
        cmd = [_nvcc()] + flags + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or out.returncode != 0:
            sys.stdout.write(out.stdout)
        if out.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    return obj


def build_library(force=False, verbose=False, extra_flags=()):
    """Compile what changed and (re)link.  Safe to call from several processes at once (one rank per GPU all
    call it at start-up): an exclusive file lock serialises them, and the up-to-date check does not depend
    on where the repository is mounted (the GPU box runs a copy under another path)."""
    import fcntl
    if os.environ.get("SPNB_NO_BUILD") and os.path.exists(LIB_PATH):
        return LIB_PATH  # use the prebuilt (variant) library as it is
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, "build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose, extra_flags)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose, extra_flags):
    flags = NVCC_FLAGS + list(extra_flags) + os.environ.get("SPNB_NVCC_EXTRA", "").split()
    if force:
        for old in os.listdir(OBJ):
            if old != "build.lock":  # the lock is held (flock) by this very call
                os.remove(os.path.join(OBJ, old))
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, flags, verbose), srcs))
    stamp = os.path.join(OBJ, "link.%s.stamp" % os.path.basename(LIB_PATH))  # one stamp per variant library
    key = " ".join(os.path.basename(o) for o in objs)  # content digests, independent of the checkout path
    if force or not os.path.exists(LIB_PATH) or not os.path.exists(stamp) or open(stamp).read() != key:
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB_PATH]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        with open(stamp, "w") as fp:
            fp.write(key)
    return LIB_PATH


def needs_build():
    return not os.path.exists(LIB_PATH)


if __name__ == "__main__":
    extra = ["-Xptxas", "-v"] if "--ptxas" in sys.argv else []
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv or bool(extra),
                        extra_flags=extra))
