"""TEST INFRASTRUCTURE: builds the CPU oracle (oracle/fgoicp_oracle.c) into oracle/libfgoicp_oracle.so.

-ffp-contract=off keeps gcc from fusing a*b+c on its own: every fused operation in the oracle is
an explicit fmaf(), placed where the reference's sm_100 SASS has an FFMA.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fgoicp_oracle.c")
OUT = os.path.join(HERE, "libfgoicp_oracle.so")


def build(force: bool = False) -> str:
    deps = [SRC, os.path.join(HERE, "svd3.h")]
    if (not force and os.path.exists(OUT)
            and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps)):
        return OUT
    cmd = ["gcc", "-std=c11", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp",
           "-fvisibility=hidden", "-shared", "-fPIC", "-Wall", "-Wextra", "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
