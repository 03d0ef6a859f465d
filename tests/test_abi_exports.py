"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares;
compute entry points refuse to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from minimd_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "minimd_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mmd_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_a_real_surface():
    syms = declared_symbols()
    assert len(syms) >= 35
    for must in ("mmd_force_lj_compute", "mmd_force_eam_compute", "mmd_neigh_build", "mmd_integrate_initial",
                 "mmd_integrate_final", "mmd_comm_borders", "mmd_run"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_every_declared_symbol():
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.mmd_abi_version() == 1
    assert isinstance(lib.mmd_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    lib = _lib.load()
    if lib.mmd_device_count() > 0:
        pytest.skip("a GPU is visible")
    h = C.c_void_p()
    rc = lib.mmd_ctx_create(0, 8, 4, None, C.byref(h))
    assert rc == 3  # MMD_ERR_NODEVICE
    assert b"no CUDA device" in lib.mmd_last_error()
    assert not h.value


def test_bad_arguments_are_status_codes_not_crashes():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.mmd_ctx_create(0, 5, 4, None, C.byref(h)) == 1
    assert lib.mmd_ctx_create(0, 8, 0, None, C.byref(h)) == 1
    assert lib.mmd_ctx_sync(None) == 1
    assert lib.mmd_atom_pbc(None) == 1
    assert lib.mmd_ctx_destroy(None) == 0
