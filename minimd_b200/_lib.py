"""ctypes loader for libminimd_b200.so (the C ABI declared in include/minimd_b200.h).

The library is built in-tree by minimd_b200.build; a missing library is a hard error -- there is
no Python/CPU fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libminimd_b200.so")

MMD_MAX_SWAPS = 32


class MmdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"minimd_b200 error {code}: {msg}")
        self.code = code


class BinGeometry(C.Structure):
    _fields_ = [("nbinx", C.c_int), ("nbiny", C.c_int), ("nbinz", C.c_int),
                ("mbinx", C.c_int), ("mbiny", C.c_int), ("mbinz", C.c_int),
                ("mbinxlo", C.c_int), ("mbinylo", C.c_int), ("mbinzlo", C.c_int),
                ("bininvx", C.c_double), ("bininvy", C.c_double), ("bininvz", C.c_double)]


class SwapTable(C.Structure):
    _fields_ = [("me", C.c_int), ("nprocs", C.c_int), ("nswap", C.c_int),
                ("need", C.c_int * 3), ("procgrid", C.c_int * 3), ("procneigh", (C.c_int * 2) * 3),
                ("sendproc", C.c_int * MMD_MAX_SWAPS), ("recvproc", C.c_int * MMD_MAX_SWAPS),
                ("pbc_any", C.c_int * MMD_MAX_SWAPS), ("pbc_flagx", C.c_int * MMD_MAX_SWAPS),
                ("pbc_flagy", C.c_int * MMD_MAX_SWAPS), ("pbc_flagz", C.c_int * MMD_MAX_SWAPS),
                ("slablo", C.c_double * MMD_MAX_SWAPS), ("slabhi", C.c_double * MMD_MAX_SWAPS)]


class RunParams(C.Structure):
    _fields_ = [("ntimes", C.c_int), ("first_step", C.c_int), ("total_steps", C.c_int),
                ("neigh_every", C.c_int), ("sort_every", C.c_int), ("thermo_nstat", C.c_int),
                ("halfneigh", C.c_int), ("ghost_newton", C.c_int), ("force_style", C.c_int),
                ("dt", C.c_double), ("dtforce", C.c_double), ("mass", C.c_double)]


class ThermoSample(C.Structure):
    _fields_ = [("step", C.c_int), ("sum_mv2", C.c_double), ("eng_vdwl", C.c_double), ("virial", C.c_double)]


# name -> (restype, argtypes); every symbol include/minimd_b200.h declares
_P = C.c_void_p
_I = C.c_int
_IP = C.POINTER(C.c_int)
_D = C.c_double
_DP = C.POINTER(C.c_double)
_LLP = C.POINTER(C.c_longlong)
SIGNATURES = {
    "mmd_last_error": (C.c_char_p, []),
    "mmd_abi_version": (_I, []),
    "mmd_device_count": (_I, []),
    "mmd_ctx_create": (_I, [_I, _I, _I, _P, C.POINTER(_P)]),
    "mmd_ctx_destroy": (_I, [_P]),
    "mmd_ctx_sync": (_I, [_P]),
    "mmd_ctx_stream": (_P, [_P]),
    "mmd_ctx_launches": (C.c_longlong, [_P]),
    "mmd_atom_set_box": (_I, [_P, _DP, _DP, _DP]),
    "mmd_atom_upload": (_I, [_P, _P, _P, _P, _I, _I]),
    "mmd_atom_split": (_I, [_P, _I]),
    "mmd_atom_update": (_I, [_P, _P, _P, _I, _I, _I]),
    "mmd_atom_download": (_I, [_P, _P, _P, _P, _P, _I, _I, _I]),
    "mmd_atom_counts": (_I, [_P, _IP, _IP, _IP]),
    "mmd_atom_pbc": (_I, [_P]),
    "mmd_atom_sort": (_I, [_P]),
    "mmd_neigh_setup": (_I, [_P, C.POINTER(BinGeometry), _P, _I, _P]),
    "mmd_neigh_binatoms": (_I, [_P, _I, _IP, _IP]),
    "mmd_neigh_build": (_I, [_P, _I, _I, _IP, _LLP]),
    "mmd_neigh_download": (_I, [_P, _P, _P, _I, _I]),
    "mmd_neigh_upload": (_I, [_P, _P, _P, _I, _I]),
    "mmd_neigh_bins_download": (_I, [_P, _P, _P, _I]),
    "mmd_neigh_atom_bins_download": (_I, [_P, _P, _I]),
    "mmd_force_lj_setup": (_I, [_P, _P, _P, _P]),
    "mmd_force_lj_compute": (_I, [_P, _I, _I, _I, _P, _P]),
    "mmd_force_eam_setup": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _D, _D, _P]),
    "mmd_force_eam_compute": (_I, [_P, _I, _I, _P, _P]),
    "mmd_integrate_initial": (_I, [_P, _D, _D]),
    "mmd_integrate_final": (_I, [_P, _D]),
    "mmd_thermo_sum_mv2": (_I, [_P, _D, _DP]),
    "mmd_comm_setup": (_I, [_P, C.POINTER(SwapTable)]),
    "mmd_comm_nccl_unique_id": (_I, [_P]),
    "mmd_comm_nccl_init": (_I, [_P, _P, _I, _I]),
    "mmd_comm_exchange": (_I, [_P]),
    "mmd_comm_borders": (_I, [_P]),
    "mmd_comm_communicate": (_I, [_P]),
    "mmd_comm_reverse_communicate": (_I, [_P]),
    "mmd_comm_swap_counts": (_I, [_P, _IP, _IP, _IP]),
    "mmd_comm_sendlist_download": (_I, [_P, _I, _P, _I]),
    "mmd_comm_allreduce": (_I, [_P, _DP, _I, _I]),
    "mmd_run": (_I, [_P, C.POINTER(RunParams), C.POINTER(ThermoSample), _I, _IP, C.POINTER(C.c_float)]),
    "mmd_run_phase_times": (_I, [_P, _DP, _LLP, _I]),
    "mmd_query_int": (_I, [_P, C.c_char_p, _LLP]),
    "mmd_set_option": (_I, [_P, C.c_char_p, C.c_longlong]),
}

_lib = None


def _preload_nccl() -> None:
    """If torch's bundled NCCL exists, map it first so one NCCL serves torch and this library."""
    if "torch" in sys.modules:
        return
    for sp in sys.path:
        hits = glob.glob(os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so.2"))
        if hits:
            try:
                C.CDLL(hits[0], mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            return


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m minimd_b200.build` (or __graft_entry__.build()); "
                              "minimd_b200 has no CPU fallback")
        _preload_nccl()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise MmdError(code, load().mmd_last_error().decode(errors="replace"))
