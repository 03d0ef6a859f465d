"""minimd_b200 -- B200 (sm_100a) implementation of the miniMD hot path behind a C ABI.

The product is the native library (minimd_b200/lib/libminimd_b200.so, declared in
include/minimd_b200.h) and the C++ host layer / driver built on it (minimd_b200/csrc/host).
This Python package only binds the C ABI for tests and bench.py; it contains no numerics.
"""
from ._lib import MmdError, load  # noqa: F401
from .api import Context, nccl_unique_id  # noqa: F401
from .host import HostError, Simulation, input_file, load_host  # noqa: F401

__all__ = ["Context", "MmdError", "load", "nccl_unique_id", "Simulation", "HostError", "input_file", "load_host"]
