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
        if f.endswith((".cu", ".cuh", ".h")):
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
    units = [f for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    objs = [os.path.join(CSRC, "." + f[:-3] + ".%d.o" % os.getpid()) for f in units]
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas", "-v"] if verbose else [])
    try:
        # translation units compile concurrently (the kernels; the CUB-based ingestion), then one device link
        procs = [subprocess.Popen([nvcc] + compile_flags + ["-c", "-o", o, os.path.join(CSRC, f)]) for f, o in zip(units, objs)]
        rcs = [p.wait() for p in procs]
        if any(rcs):
            raise subprocess.CalledProcessError(max(rcs), "nvcc -c")
        subprocess.run([nvcc] + [f for f in NVCC_FLAGS if f not in ("-lineinfo", "-O3", "-std=c++17")] + ["-o", tmp] + objs, check=True)
        os.replace(tmp, LIB)
    finally:
        for f in objs + [tmp]:
            if os.path.exists(f):
                os.remove(f)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
