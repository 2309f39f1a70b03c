"""Build the CUDA engine in-tree: abides_b200/csrc/libabides_b200.so (sm_100a only)."""
import os
import subprocess
import sys

_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(_DIR), "include")
OUT = os.path.join(CSRC, "libabides_b200.so")
SOURCES = ["engine.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                      # no implicit contraction: see include/abides_dd_math.h
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "--use_fast_math=false",
]


def build_cuda(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(INCLUDE, h) for h in ("abides_b200.h", "abides_dd_math.h", "abides_math_tables.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-o", OUT] + srcs
    print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build_cuda(force=True, verbose="-v" in sys.argv)
