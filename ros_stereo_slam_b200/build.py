"""Builds libvo_b200.so (hand-written CUDA for sm_100a behind the C ABI of include/vo_b200.h).

    python -m ros_stereo_slam_b200.build [--force]

nvcc cross-compiles without a GPU.  The library is built IN-TREE
(ros_stereo_slam_b200/libvo_b200.so) so that it travels to the GPU box with the repo
snapshot; it is git-ignored.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvo_b200.so")
OBJ = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# VO_HELPERS_NOINLINE: keep the solver helpers as separate device functions (small, robust
# compile units; see selfcheck.cu for why the code shape matters).
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-DVO_HELPERS_NOINLINE", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "-Xcompiler", "-O2", "--expt-relaxed-constexpr"]
# FP64 solver files follow OpenCV's operation order: no FMA contraction anywhere in them.
SOURCES = {
    "api.cu": ["--fmad=false"],
    "pyramid.cu": [],
    "lk.cu": [],
    "points.cu": ["--fmad=false"],
    "ransac.cu": ["--fmad=false"],
    "refine.cu": ["--fmad=false"],
    "selfcheck.cu": ["--fmad=false"],
    "sor.cu": ["--fmad=false"],
    "sgbm.cu": ["--fmad=false"],
    "orb.cu": ["--fmad=false"],
    "aux.cu": [],
}
HEADERS = ["common.cuh", "cvmath.cuh", "fmat7.cuh", "jacobi_warp.cuh", "orb_pattern.h", os.path.join("..", "..", "include", "vo_b200.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs = []
    objs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            logs = list(ex.map(run, jobs))
        if verbose:
            for l in logs:
                sys.stderr.write(l)
    if force or jobs or _stale(OUT, objs):
        # the symbols of include/vo_b200.h are exported explicitly (see VO_API in api.cu)
        cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-Xcompiler", "-fPIC", "-lcudart"]
        run(cmd)
    return OUT


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(out)
