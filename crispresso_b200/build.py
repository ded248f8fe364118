"""Build libcrgpu.so in-tree with nvcc for sm_100a (no other target, no JIT cache).

Each translation unit is compiled to its own object (in parallel, re-used while the source and every
header are older than it), then linked; `--force` rebuilds everything.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SOURCES = ["gotoh_fill.cu", "gotoh_score2.cu", "traceback_walk.cu", "quantify.cu", "int_peak.cu", "crgpu_api.cu", "hotpath.cu", "alleles.cu",
           "flash_merge.cu", "fastq_index.cu"]
LIB = os.path.join(HERE, "libcrgpu.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-shared", "-cudart", "shared"]


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + \
           [os.path.join(HERE, "..", "include", "crgpu.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    return _stale(LIB, srcs + _headers())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else [])
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(s):
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or extra or _stale(obj, [src] + hdrs):
            subprocess.check_call([nvcc] + NVCC_FLAGS + extra + ["-c", "-o", obj, src])
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    subprocess.check_call([nvcc] + NVCC_FLAGS + LINK_FLAGS + ["-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
