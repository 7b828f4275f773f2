"""Builds the CUDA library of the B200 Go-ICP hot path in-tree: fast_go_icp_b200/libfgoicp_b200.so.

sm_100a only (-gencode arch=compute_100a,code=sm_100a), -lineinfo so ncu's source page maps to
these files.  nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU
box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "csrc", "_obj")
OUT = os.path.join(HERE, "libfgoicp_b200.so")

CU_SOURCES = ["ctx.cu", "bounds.cu", "bounds_phased.cu", "nn_icp.cu", "bnb.cu", "probe.cu", "preprocess.cu"]
CPP_SOURCES = ["fgoicp_host.cpp"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-Xptxas", "-v",
              "-I" + os.path.join(ROOT, "include")]


def _newer(out, deps):
    return os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps if os.path.exists(d))


def _compile(src):
    obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".h"))]
    hdrs += [os.path.join(ROOT, "include", "fgoicp_c.h")]
    hdrs += [os.path.join(ROOT, "include", "fgoicp", f) for f in os.listdir(os.path.join(ROOT, "include", "fgoicp"))]
    hdrs += [os.path.join(ROOT, "include", "glm", f) for f in os.listdir(os.path.join(ROOT, "include", "glm"))]
    if _newer(obj, [path] + hdrs):
        return obj, ""
    cmd = ["nvcc"] + NVCC_FLAGS + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = [s for s in CU_SOURCES + CPP_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if force:
        for s in srcs:
            o = os.path.join(OBJDIR, os.path.splitext(s)[0] + ".o")
            if os.path.exists(o):
                os.remove(o)
    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        results = list(ex.map(_compile, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if not _newer(OUT, objs):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-lcudart", "-ldl"]
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
