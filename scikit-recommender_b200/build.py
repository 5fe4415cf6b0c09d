"""Build the C-ABI CUDA library in-tree: scikit-recommender_b200/libskrec_b200.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libskrec_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    out = [os.path.join(HERE, "..", "include", "skrec_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh")):
            out.append(os.path.join(CSRC, f))
    return out


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB):
        t = os.path.getmtime(LIB)
        if all(os.path.getmtime(s) <= t for s in sources()):
            return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("SKR_NVCC_EXTRA", "").split()  # e.g. -DSKR_TC_TRACE=1 for tools/trace_tiles.py
    tmp = LIB + ".tmp.%d" % os.getpid()  # built next to the target and renamed: a snapshot never sees a half-written library
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, os.path.join(CSRC, "skrec_b200.cu")]
    try:
        subprocess.run(cmd, check=True)
        os.replace(tmp, LIB)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
