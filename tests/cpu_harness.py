"""TEST INFRASTRUCTURE: builds build/fgoicp_harness_cpu -- tests/cpp/fgoicp_harness.cpp plus the C++ host driver
(fast_go_icp_b200/csrc/fgoicp_host.cpp, unchanged) linked against tests/cpp/oracle_abi.c, an oracle-backed stand-in for
the C ABI, instead of the CUDA library -- so that the drop-in class can be exercised without a GPU
(tests/test_cpp_host_cpu.py).  Lives under tests/ because it links the oracle; the package never does."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "fgoicp_harness.cpp")


def build_cpu(force: bool = False):
    out = os.path.join(ROOT, "build", "fgoicp_harness_cpu")
    host = os.path.join(ROOT, "fast_go_icp_b200", "csrc", "fgoicp_host.cpp")
    abi = os.path.join(ROOT, "tests", "cpp", "oracle_abi.c")
    oracle_dir = os.path.join(ROOT, "oracle")
    deps = [SRC, host, abi, os.path.join(oracle_dir, "libfgoicp_oracle.so"), os.path.join(ROOT, "include", "fgoicp", "fgoicp.hpp")]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(d) for d in deps):
        return out
    obj = os.path.join(ROOT, "build", "oracle_abi.o")
    subprocess.run(["gcc", "-std=c11", "-O2", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), "-c", abi, "-o", obj], check=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), SRC, host, obj, "-o", out,
                    "-L" + oracle_dir, "-lfgoicp_oracle", "-Wl,-rpath," + oracle_dir, "-Wl,-rpath,$ORIGIN/../oracle", "-lpthread"], check=True)
    return out


if __name__ == "__main__":
    print(build_cpu(force=True))
