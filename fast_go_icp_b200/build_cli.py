"""Builds the reference's own command-line program (src/main.cpp + src/utilities.hpp, UNCHANGED, read
from /root/reference) against this repository's headers (include/fgoicp, include/glm) and library.
This is the drop-in demonstration: the caller code is untouched, only the library beneath it changed.

Output: build/fast-go-icp (git-ignored; travels to the GPU box).  No-op when the reference tree is absent.
-fpermissive is needed because utilities.hpp:38,61 qualify member declarations with the class name,
which GCC rejects by default (MSVC accepts it)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "build", "fast-go-icp")


def build(force: bool = False):
    main_cpp = os.path.join(REF, "src", "main.cpp")
    if not os.path.exists(main_cpp):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    lib = os.path.join(HERE, "libfgoicp_b200.so")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(lib):
        return OUT
    cmd = ["g++", "-std=c++17", "-O2", "-fpermissive", "-w",
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(REF, "src"),
           "-I" + os.path.join(REF, "external", "include"),
           main_cpp, "-o", OUT, "-L" + HERE, "-lfgoicp_b200", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/../fast_go_icp_b200"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
