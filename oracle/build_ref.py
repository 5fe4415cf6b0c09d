"""Build recipe for the oracle (test infrastructure; see oracle/skr_oracle.c header).

Two products, both git-ignored but shipped to the GPU box by gpurun:

  oracle/_build/libskr_oracle.so   gcc build of oracle/skr_oracle.c (the CPU restatement).
  oracle/_ref/                     the UNMODIFIED reference, compiled from the sources where
                                   they lie under /root/reference -- binaries only:
      libref_eval.so               oracle/ref_shim.cpp + reference evaluate.h/metric.h/thread_pool.h
      refpkg/cython/pyx_*.so       Cython builds of skrec/utils/py/cython/pyx_{utils,eval_matrix,sort}.pyx
      refpkg/{evaluator,batch_iterator}.so
                                   Cython builds of skrec/utils/py/{evaluator,batch_iterator}.py
                                   (so the reference's own RankingEvaluator.evaluate runs on the GPU
                                   box, where /root/reference does not exist)
      refpkg/__init__.py, refpkg/cython/__init__.py, colorama/__init__.py
                                   generated glue: package markers and an 8-line colorama stub
                                   (colorama is not installed; evaluator.py:10 imports it).

The reference's own build (setup.py:58-72) cythonizes with -std=c++11 only and distutils adds
-O2; the same flags are used here.  `python oracle/build_ref.py` is idempotent.  When
/root/reference is absent (GPU box) the _ref step is skipped and the prebuilt files are used.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SKR_REFERENCE_ROOT", "/root/reference")
REF_PY = os.path.join(REF_ROOT, "skrec", "utils", "py")
REF_CY = os.path.join(REF_PY, "cython")
OUT_REF = os.path.join(HERE, "_ref")
OUT_BUILD = os.path.join(HERE, "_build")


def _run(cmd):
    subprocess.run(cmd, check=True)


def _newer(src_list, dst):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in src_list)


def build_oracle(force=False):
    os.makedirs(OUT_BUILD, exist_ok=True)
    src = os.path.join(HERE, "skr_oracle.c")
    dst = os.path.join(OUT_BUILD, "libskr_oracle.so")
    if force or _newer([src], dst):
        _run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", src, "-o", dst, "-lm"])
    return dst


def _ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX")


def _cython_build(src, modname, dst_dir, tmp, cplus):
    """cythonize `src` where it lies; only the .so lands in dst_dir."""
    import numpy

    ext = ".cpp" if cplus else ".c"
    gen = os.path.join(tmp, modname + ext)
    cmd = [sys.executable, "-m", "cython", "-3", src, "-o", gen]
    if cplus:
        cmd.insert(4, "--cplus")
    else:
        # plain .py modules: keep Python semantics (annotations stay hints, e.g. an OrderedDict is a
        # fine `Dict`), so the compiled evaluator behaves exactly like the interpreted one
        cmd[4:4] = ["-X", "annotation_typing=False"]
    _run(cmd)
    out = os.path.join(dst_dir, modname + _ext_suffix())
    cc = ["g++", "-std=c++11"] if cplus else ["gcc"]
    _run(cc + ["-O2", "-fPIC", "-shared", "-pthread", "-fwrapv", "-w",
               "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
               "-I" + REF_CY, "-I" + os.path.join(REF_CY, "include"), gen, "-o", out])
    return out


def build_ref(force=False):
    if not os.path.isdir(REF_CY):
        return None  # GPU box: use what travelled
    stamp = os.path.join(OUT_REF, ".built")
    srcs = [os.path.join(REF_CY, f) for f in ("pyx_utils.pyx", "pyx_eval_matrix.pyx", "pyx_sort.pyx")] + \
           [os.path.join(REF_CY, "include", f) for f in ("evaluate.h", "metric.h", "thread_pool.h")] + \
           [os.path.join(REF_PY, f) for f in ("evaluator.py", "batch_iterator.py")] + \
           [os.path.join(HERE, "ref_shim.cpp"), os.path.abspath(__file__)]
    if not force and not _newer(srcs, stamp):
        return OUT_REF
    pkg = os.path.join(OUT_REF, "refpkg")
    cy = os.path.join(pkg, "cython")
    tmp = os.path.join(OUT_REF, "_tmp")
    for d in (cy, tmp, os.path.join(OUT_REF, "colorama")):
        os.makedirs(d, exist_ok=True)
    # 1. reference C++ headers behind a C doorway
    _run(["g++", "-O2", "-std=c++11", "-fPIC", "-shared", "-pthread", "-w",
          "-I" + os.path.join(REF_CY, "include"), os.path.join(HERE, "ref_shim.cpp"),
          "-o", os.path.join(OUT_REF, "libref_eval.so")])
    # 2. reference Cython modules and the reference evaluator, as extension modules
    _cython_build(os.path.join(REF_CY, "pyx_utils.pyx"), "pyx_utils", cy, tmp, True)
    _cython_build(os.path.join(REF_CY, "pyx_eval_matrix.pyx"), "pyx_eval_matrix", cy, tmp, True)
    _cython_build(os.path.join(REF_CY, "pyx_sort.pyx"), "pyx_sort", cy, tmp, True)  # pins skrec_b200.sort
    _cython_build(os.path.join(REF_PY, "batch_iterator.py"), "batch_iterator", pkg, tmp, False)
    _cython_build(os.path.join(REF_PY, "evaluator.py"), "evaluator", pkg, tmp, False)
    shutil.rmtree(tmp, ignore_errors=True)
    # 3. generated glue (not reference code)
    with open(os.path.join(pkg, "__init__.py"), "w") as f:
        f.write("# generated by oracle/build_ref.py\n")
    with open(os.path.join(cy, "__init__.py"), "w") as f:
        f.write("# generated by oracle/build_ref.py\nfrom .pyx_eval_matrix import eval_score_matrix\n")
    with open(os.path.join(OUT_REF, "colorama", "__init__.py"), "w") as f:
        f.write("# generated by oracle/build_ref.py: colour codes become empty strings\n"
                "class _Empty(object):\n"
                "    def __getattr__(self, name):\n"
                "        return ''\n"
                "Fore = _Empty()\nBack = _Empty()\nStyle = _Empty()\n"
                "def init(*args, **kwargs):\n    pass\n")
    with open(stamp, "w") as f:
        f.write("ok\n")
    return OUT_REF


if __name__ == "__main__":
    force = "--force" in sys.argv
    print("oracle:", build_oracle(force))
    print("reference:", build_ref(force) or "skipped (no %s); using prebuilt oracle/_ref" % REF_ROOT)
