"""hpmpc_b200 -- B200-native batched Riccati / box-IPM engine behind HPMPC's C API.

The product is the shared library hpmpc_b200/lib/libhpmpc_b200.so (C host shim + sm_100a kernels, built by
hpmpc_b200/csrc/Makefile).  The Python modules here are harness code for tests and benchmarks:
  capi      ctypes access to the library (and to any other library exporting HPMPC's symbols)
  problems  the reference's mass-spring test problems
"""
from . import problems  # noqa: F401
