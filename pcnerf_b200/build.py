"""Build libpcnerf_b200.so (C ABI, include/pcnerf_b200.h) in-tree with nvcc for sm_100a.

    python -m pcnerf_b200.build [--force]

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot (it is git-ignored).
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libpcnerf_b200.so")

SOURCES = {
    # file: extra flags.  aabb.cu reproduces numpy fp64 arithmetic -> no FMA contraction anywhere in that file.
    "error.cu": [],
    "aabb.cu": ["-fmad=false"],
    "sample_encode.cu": [],
    "composite.cu": [],
    "search.cu": [],
    "mlp.cu": [],
    "mlp_tc.cu": [],
    "affine.cu": [],
    "affine_rays.cu": [],
    "metrics.cu": ["-fmad=false"],
    "frame.cu": ["-fmad=false"],
    "optim.cu": [],
    "loss.cu": [],
}
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
          "-Xcompiler", "-fPIC"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "mlp_layout.h"), os.path.join(CSRC, "mlp_small.cuh"), os.path.join(CSRC, "encode.cuh"),
           os.path.join(os.path.dirname(HERE), "include", "pcnerf_b200.h")]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + HEADERS):
            jobs.append(([nvcc] + COMMON + extra + ["-c", s, "-o", o], src))

    def run(job):
        cmd, name = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (name, r.stdout, r.stderr))
        if verbose:
            print("compiled", name)
        return name

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("linked", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
