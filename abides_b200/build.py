"""Build the CUDA engine in-tree: abides_b200/csrc/libabides_b200.so (sm_100a only)."""
import os
import subprocess
import sys

_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(_DIR), "include")
OUT = os.path.join(CSRC, "libabides_b200.so")
OUT_PROF = os.path.join(CSRC, "libabides_b200_prof.so")     # -DABIDES_PROFILE: per-phase cycle counters (diagnosis only)
SOURCES = ["engine.cu", "lane_kernels.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                      # no implicit contraction: see include/abides_dd_math.h
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "--use_fast_math=false",
]


def build_cuda(force=False, verbose=False, profile=False, cta_warps=None, time_slices=None, lane_cta=None, defines=(), tag=None):
    """Compile csrc/engine.cu for sm_100a.  profile=True builds the instrumented variant next to it;
    cta_warps=W builds libabides_b200_w<W>.so with W instances per CTA (experiments; the default is 1)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in ("engine_device.inc", "lane_device.inc", "lane_types.h", "lane_variants.inc", "lane_host_common.h", "lane_kernels.h", "rng_setup_device.inc")] + [os.path.join(INCLUDE, h) for h in ("abides_b200.h", "abides_dd_math.h", "abides_math_tables.h")]
    out = OUT_PROF if profile else OUT
    if cta_warps:
        out = out.replace(".so", "_w%d.so" % cta_warps)
    if time_slices:
        out = out.replace(".so", "_q%d.so" % time_slices)
    if lane_cta:
        out = out.replace(".so", "_l%d.so" % lane_cta)
    if tag:
        out = out.replace(".so", "_%s.so" % tag)
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + (["-DABIDES_PROFILE"] if profile else [])
    if cta_warps:
        flags.append("-DABM_CTA_WARPS=%d" % cta_warps)
    if time_slices:
        flags.append("-DABM_TIME_SLICES=%d" % time_slices)
    if lane_cta:
        flags.append("-DLANE_CTA_THREADS=%d" % lane_cta)
    flags += ["-D" + d for d in defines]
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-o", out] + srcs
    print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    w = [a for a in sys.argv if a.startswith("--cta-warps=")]
    build_cuda(force=True, verbose="-v" in sys.argv, profile="--profile" in sys.argv,
               cta_warps=int(w[0].split("=")[1]) if w else None,
               time_slices=next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--time-slices=")), None),
               lane_cta=next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--lane-cta=")), None),
               defines=[a.split("=", 1)[1] for a in sys.argv if a.startswith("--define=")],
               tag=next((a.split("=")[1] for a in sys.argv if a.startswith("--tag=")), None))
