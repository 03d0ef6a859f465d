#!/usr/bin/env python
"""Regenerate the committed golden fixtures.  Runs ONLY in the build container, where the
reference tree (/root/reference) and the reference binaries (oracle/_ref/, oracle/build_ref.sh)
exist.  Nothing in tests/ or bench.py needs /root/reference at run time -- they read the JSON /
NPZ files this script writes.

  reference_logs.json   thermo lines (step, T, U, P) + natoms + PERF_SUMMARY of every log the
                        reference ships under tests/reference_output/ (16 MPI ranks, FP64).
  reference_runs.json   the unmodified reference binary run here on the BASELINE.json configs and
                        on small parity cases: 10-digit T/U/P (YAML), nlocal, nghost, sum(numneigh).
  cu_u6_funcfl.npz      the numbers of the reference's Cu EAM funcfl table (ref/Cu_u6.eam) as
                        float64 arrays; minimd_b200.host.write_funcfl() turns them back into a
                        funcfl text file for runs (17 significant digits => identical doubles).

usage: python tests/golden/make_golden.py [--big]     (--big adds -s 80 / EAM -s 64; minutes)
       python tests/golden/make_golden.py --weak       (only ADDS the weak-scaling boxes of bench.py --gpus 2/4/8:
                                                        80x80x160, 80x160x160, 160^3 cells; ~10 minutes, ~10 GB)
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("MINIMD_REFERENCE", "/root/reference")

from oracle.oracle import Config, run_reference  # noqa: E402


def parse_log(path):
    txt = open(path).read()
    out = {"steps": [], "T": [], "U": [], "P": []}
    m = re.search(r"# Atoms: (\d+)", txt)
    out["natoms"] = int(m.group(1))
    m = re.search(r"unit cells: (\d+) (\d+) (\d+)", txt)
    out["cells"] = [int(m.group(i)) for i in (1, 2, 3)]
    m = re.search(r"# MPI processes: (\d+)", txt)
    out["mpi_ranks"] = int(m.group(1))
    body = txt.split("# Timestep T U P Time")[1].split("# Performance Summary")[0]
    for ln in body.strip().splitlines():
        p = ln.split()
        if len(p) == 5:
            out["steps"].append(int(p[0]))
            out["T"].append(float(p[1]))
            out["U"].append(float(p[2]))
            out["P"].append(float(p[3]))
    m = re.search(r"#?(\d+) (\d+) (\d+) (\d+) (\S+) (\S+) (\S+) (\S+) (\S+) (\S+) (\S+) PERF_SUMMARY", txt)
    if m:
        out["perf_atom_steps_per_s"] = float(m.group(10))
    return out


def shipped_logs():
    d = os.path.join(REF, "tests", "reference_output")
    logs = {}
    for fn in sorted(os.listdir(d)):
        if re.match(r"\d+k\.(lj|eam)(-data)?$", fn):
            logs[fn] = parse_log(os.path.join(d, fn))
            logs[fn]["source"] = f"tests/reference_output/{fn}"
    return logs


def reference_runs(big, weak_only=False):
    cases = {}

    def add(name, cfg, precision="f64", threads=1):
        r = run_reference(cfg, precision, threads=threads)
        c = cfg.resolved()
        cases[name] = {
            "config": {k: getattr(c, k) for k in ("nx", "ny", "nz", "ntimes", "force", "halfneigh", "ghost_newton",
                                                  "neigh_every", "sort", "thermo_nstat", "nbins", "ntypes")},
            "precision": precision, "threads": threads,
            "steps": r.steps, "T": r.T, "U": r.U, "P": r.P,
            "nlocal": r.nlocal, "nghost": r.nghost, "neighs": r.neighs, "natoms": r.natoms,
        }
        print(name, r.steps[-1], r.T[-1], r.U[-1], r.P[-1], r.nghost, r.neighs, flush=True)

    if weak_only:  # 80^3 cells per GPU on 2 / 4 / 8 ranks
        add("lj_80x80x160_half", Config(nx=80, ny=80, nz=160), threads=8)
        add("lj_80x160x160_half", Config(nx=80, ny=160, nz=160), threads=8)
        add("lj_s160_half", Config(nx=160, ny=160, nz=160), threads=8)
        return cases
    for force in ("lj", "eam"):
        for half, gn in ((1, 1), (1, 0), (0, 0)):
            add(f"{force}_s8_half{half}_gn{gn}", Config(nx=8, ny=8, nz=8, force=force, halfneigh=half, ghost_newton=gn,
                                                          thermo_nstat=10))
    add("lj_s8_every_step", Config(nx=8, ny=8, nz=8, thermo_nstat=1))
    add("lj_s10", Config(nx=10, ny=10, nz=10, ntimes=1000, thermo_nstat=100))
    add("eam_s10", Config(nx=10, ny=10, nz=10, ntimes=300, force="eam", thermo_nstat=100))
    add("lj_6x8x10_nosort", Config(nx=6, ny=8, nz=10, ntimes=60, thermo_nstat=20, sort=0))
    add("lj_s32", Config(nx=32, ny=32, nz=32, thermo_nstat=100), threads=8)
    add("lj_s8_f32_full", Config(nx=8, ny=8, nz=8, halfneigh=0, ghost_newton=0, thermo_nstat=10), precision="f32")
    if big:
        add("lj_s80_half", Config(nx=80, ny=80, nz=80), threads=8)
        add("lj_s80_full_gn0", Config(nx=80, ny=80, nz=80, halfneigh=0, ghost_newton=0), threads=8)
        add("eam_s64_full", Config(nx=64, ny=64, nz=64, force="eam", halfneigh=0, ghost_newton=0), threads=8)
    return cases


def funcfl_table():
    """Numbers of ref/Cu_u6.eam (format: ref/force_eam.cpp:505-582)."""
    with open(os.path.join(REF, "ref", "Cu_u6.eam")) as fh:
        fh.readline()
        l2 = fh.readline().split()
        l3 = fh.readline().split()
        vals = np.array(fh.read().split(), dtype=np.float64)
    nrho, drho, nr, dr, cut = int(l3[0]), float(l3[1]), int(l3[2]), float(l3[3]), float(l3[4])
    assert vals.size == nrho + 2 * nr
    return dict(atomic_number=int(l2[0]), mass=float(l2[1]), lattice=float(l2[2]), nrho=nrho, drho=drho, nr=nr,
                dr=dr, cut=cut, frho=vals[:nrho], zr=vals[nrho:nrho + nr], rhor=vals[nrho + nr:])


def main():
    big = "--big" in sys.argv
    if "--weak" in sys.argv:
        path = os.path.join(HERE, "reference_runs.json")
        runs = json.load(open(path))
        runs.update(reference_runs(False, weak_only=True))
        with open(path, "w") as fh:
            json.dump(runs, fh, indent=0)
        return
    with open(os.path.join(HERE, "reference_logs.json"), "w") as fh:
        json.dump(shipped_logs(), fh, indent=0)
    np.savez_compressed(os.path.join(HERE, "cu_u6_funcfl.npz"), **funcfl_table())
    path = os.path.join(HERE, "reference_runs.json")
    runs = reference_runs(big)
    if not big and os.path.exists(path):  # keep previously generated big cases
        old = json.load(open(path))
        for k, v in old.items():
            runs.setdefault(k, v)
    with open(path, "w") as fh:
        json.dump(runs, fh, indent=0)


if __name__ == "__main__":
    main()
