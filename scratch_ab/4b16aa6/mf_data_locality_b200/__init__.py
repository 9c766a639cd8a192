"""mf_data_locality_b200 -- B200-native (sm_100a, FP64) CEED BP4 CG hot path behind the
operator/solver surface of peterrum/mf_data_locality.

  csrc/   CUDA kernels + the C ABI of include/bp4.h        -> libbp4.so
  host/   C++ mirror of the reference's host-side surface  -> libbp4_host.so, bench executables
  capi.py ctypes binding of the C ABI (tests / bench / smoke)
"""
from . import capi  # noqa: F401
