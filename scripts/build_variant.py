"""Builds an experimental variant of the library: one source recompiled with extra -D flags, linked with the
regular objects into build/variants/lib_<name>.so (select it at run time with FGOICP_LIB=<path>).
usage: python scripts/build_variant.py <name> <source.cu> "<extra nvcc flags>" """
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fast_go_icp_b200 import build as B
name, src, extra = sys.argv[1], sys.argv[2], sys.argv[3].split()
B.build()
vdir = os.path.join(ROOT, "build", "variants")
os.makedirs(vdir, exist_ok=True)
if src == "all":
    objs = []
    for s_ in B.CU_SOURCES + B.CPP_SOURCES:
        o = os.path.join(vdir, "%s_%s.o" % (os.path.splitext(s_)[0], name))
        cmd = ["nvcc"] + B.NVCC_FLAGS + extra + (["-x", "cu"] if s_.endswith(".cpp") else []) + ["-c", os.path.join(B.CSRC, s_), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.exit(r.stderr)
        objs.append(o)
    out = os.path.join(vdir, "lib_%s.so" % name)
    subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-lcudart", "-ldl"], check=True)
    print(out)
    sys.exit(0)
obj = os.path.join(vdir, "%s_%s.o" % (os.path.splitext(src)[0], name))
r = subprocess.run(["nvcc"] + B.NVCC_FLAGS + extra + ["-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stderr)
for ln in r.stderr.splitlines():
    if "k_bounds_phased" in ln or "k_bnb_r3ILi1" in ln:
        print(ln[:120])
    elif "registers" in ln and prev_hit:
        print("   ", ln.strip(), "|", stack.strip())
    prev_hit = ("k_bounds_phased" in ln or "k_bnb_r3ILi1" in ln) and "Function properties" in ln or ("stack frame" in ln and prev_hit if 'prev_hit' in dir() else False)
    if "stack frame" in ln:
        stack = ln
objs = [os.path.join(B.OBJDIR, os.path.splitext(s)[0] + ".o") for s in B.CU_SOURCES + B.CPP_SOURCES if s != src] + [obj]
out = os.path.join(vdir, "lib_%s.so" % name)
subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-lcudart", "-ldl"], check=True)
print(out)
