"""In-tree build of the native artefacts (no JIT cache: the .so/.exe files travel to the GPU box).

  minimd_b200/lib/libminimd_b200.so        CUDA kernels + C ABI (include/minimd_b200.h), sm_100a
  minimd_b200/lib/libminimd_host_f64.so    C++ host classes (Atom/Neighbor/Force/Integrate/Comm/Thermo
  minimd_b200/lib/libminimd_host_f32.so      mirroring ref/) + the mmd_sim_* embedding API
  minimd_b200/bin/miniMD_b200_f64|_f32     drop-in driver executables (same CLI / input / output as ref/)

`python -m minimd_b200.build` or `minimd_b200.build.build_all()`.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib")
BIN = os.path.join(PKG, "bin")
INC = os.path.join(ROOT, "include")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = os.environ.get("CXX_HOST") or shutil.which("g++") or "g++"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"build failed: {cmd[0]} ... {cmd[-1]}")


def _nccl_flags() -> list[str]:
    if os.environ.get("MMD_NO_NCCL"):
        return []
    if os.path.exists("/usr/include/nccl.h"):
        return ["-DMMD_WITH_NCCL", "-lnccl"]
    try:  # torch-bundled NCCL
        import nvidia.nccl as n  # type: ignore
        base = os.path.dirname(n.__file__)
        return ["-DMMD_WITH_NCCL", f"-I{base}/include", f"-L{base}/lib", "-l:libnccl.so.2"]
    except Exception:
        return []


def build_device(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libminimd_b200.so")
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(INC, "minimd_b200.h"))
    if force or _newer(out, srcs):
        cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
               "-o", out, os.path.join(CSRC, "mmd_device.cu"), *_nccl_flags()]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        if os.environ.get("MMD_KERNEL_PROFILE"):   # diagnostic build: clock64 sums in the LJ tile force kernels
            cmd.insert(1, "-DMMD_KERNEL_PROFILE")
        _run(cmd)
    return out


def build_host(force: bool = False) -> list[str]:
    """C++ host layer (two precisions) + driver executables, linked against libminimd_b200.so."""
    hostdir = os.path.join(CSRC, "host")
    if not os.path.isdir(hostdir):
        return []
    os.makedirs(BIN, exist_ok=True)
    srcs = sorted(os.path.join(hostdir, f) for f in os.listdir(hostdir) if f.endswith(".cpp"))
    hdrs = [os.path.join(hostdir, f) for f in os.listdir(hostdir) if f.endswith(".h")] + [os.path.join(INC, "minimd_b200.h")]
    lib_srcs = [s for s in srcs if not s.endswith("ljs.cpp")]
    outs = []
    for prec, tag in ((2, "f64"), (1, "f32")):
        so = os.path.join(LIB, f"libminimd_host_{tag}.so")
        exe = os.path.join(BIN, f"miniMD_b200_{tag}")
        common = [CXX, "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", f"-DPRECISION={prec}", f"-I{INC}", f"-I{hostdir}", "-Wall",
                  "-Wno-unused-result"]
        link = [f"-L{LIB}", "-lminimd_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../lib"]
        if force or _newer(so, lib_srcs + hdrs):
            _run(common + ["-shared", "-o", so, *lib_srcs, *link])
        if force or _newer(exe, srcs + hdrs):
            _run(common + ["-o", exe, *srcs, *link])
        outs += [so, exe]
    return outs


def build_all(force: bool = False) -> None:
    build_device(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", *sorted(os.listdir(LIB)))
