"""Builds build/fgoicp_harness from tests/cpp/fgoicp_harness.cpp: a plain g++ program (no CUDA headers) that uses
the drop-in C++ class icp::FastGoICP exactly like the reference's src/main.cpp:46-53 does.  The GPU tests run it to
check the C++ host driver (csrc/fgoicp_host.cpp) against the Python mirror and the oracle.  Unlike build/fast-go-icp
it does not need the reference tree."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(ROOT, "tests", "cpp", "fgoicp_harness.cpp")
OUT = os.path.join(ROOT, "build", "fgoicp_harness")


def build(force: bool = False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    lib = os.path.join(HERE, "libfgoicp_b200.so")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(lib), os.path.getmtime(SRC)):
        return OUT
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), SRC, "-o", OUT,
           "-L" + HERE, "-lfgoicp_b200", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/../fast_go_icp_b200"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))

