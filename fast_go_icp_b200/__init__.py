"""fast_go_icp_b200 -- B200-native (sm_100a) data-parallel hot path of Go-ICP registration.

Layout
  csrc/      hand-written CUDA kernels + the C ABI (include/fgoicp_c.h) + the C++ host driver
             icp::FastGoICP (include/fgoicp/fgoicp.hpp), built in-tree into libfgoicp_b200.so
  capi.py    ctypes binding of the C ABI
  driver.py  host-side mirror of icp::FastGoICP with the rotation frontier sharded across ranks
  build.py   nvcc build (sm_100a only)

There is no CPU fallback anywhere in this package.
"""
from . import capi  # noqa: F401
from .capi import Context, FgoicpError  # noqa: F401
from .driver import FastGoICP  # noqa: F401
