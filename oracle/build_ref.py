"""TEST INFRASTRUCTURE: compiles the UNMODIFIED reference sources where they lie under
/root/reference into oracle/_ref/libfgoicp_ref.so (git-ignored; travels to the GPU box).

The reference needs GLM, Eigen3, CUDA and cmake (fgoicp/CMakeLists.txt:21, fgoicp/common.hpp:12-13);
GLM and Eigen are not installed here.  Instead of its build system, nvcc is run directly on its four
sources with two header stand-ins on the include path -- include/glm (the same from-scratch GLM
stand-in the drop-in boundary ships) and oracle/ref_shim/Eigen (3x3 JacobiSVD stand-in) -- plus two
forced includes (<cfloat>, <algorithm>) that the sources rely on transitively under MSVC.  No
reference source is copied or edited.  Built for plain sm_100 (the reference's own
CUDA_ARCHITECTURES all-major, fgoicp/CMakeLists.txt:33), without -use_fast_math, like the reference.

Nothing here is product code; if /root/reference is absent (GPU box) this is a no-op.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUTDIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUTDIR, "libfgoicp_ref.so")
SOURCES = ["fgoicp/registration.cu", "fgoicp/icp3d.cu", "fgoicp/fgoicp.cpp", "fgoicp/common.cpp"]


def build(force: bool = False):
    if not os.path.isdir(os.path.join(REF, "fgoicp")):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(OUTDIR, exist_ok=True)
    shim = os.path.join(HERE, "ref_shim", "ref_capi.cu")
    deps = [os.path.join(REF, s) for s in SOURCES] + [shim, os.path.join(HERE, "svd3.h"),
                                                      os.path.join(HERE, "ref_shim", "Eigen", "Dense")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    flags = ["-std=c++17", "-O3", "-gencode", "arch=compute_100,code=sm_100", "-lineinfo", "-w",
             "-Xcompiler", "-fPIC", "-include", "cfloat", "-include", "algorithm",
             "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "ref_shim"),
             "-I" + os.path.join(REF, "fgoicp")]
    objs = []
    for s in SOURCES + [shim]:
        src = s if os.path.isabs(s) else os.path.join(REF, s)
        obj = os.path.join(OUTDIR, os.path.basename(s).rsplit(".", 1)[0] + ".o")
        subprocess.run(["nvcc"] + flags + ["-x", "cu", "-c", src, "-o", obj], check=True)
        objs.append(obj)
    subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100,code=sm_100", "-o", OUT] + objs + ["-lcudart"],
                   check=True)
    for o in objs:
        os.remove(o)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
