#!/usr/bin/env python
"""Golden run of the UNMODIFIED reference binary (oracle/_ref) started from a LAMMPS data file
(ref/setup.cpp:215-301).  The data file is regenerated deterministically by tests/helpers.py, so only
the reference's answers are committed: tests/golden/reference_datafile.json."""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import lattice_for_datafile, write_lammps_data  # noqa: E402
from oracle.oracle import REF_DIR, Config, parse_reference_output, ref_binary  # noqa: E402

out = {}
for name, half, gn in (("lj_data_half", 1, 1), ("lj_data_full", 0, 0)):
    x, v, prd = lattice_for_datafile()
    with tempfile.TemporaryDirectory() as td:
        data = write_lammps_data(os.path.join(td, "atoms.data"), x, v, prd)
        cfg = Config(nx=6, ny=6, nz=6, ntimes=100, halfneigh=half, ghost_newton=gn, thermo_nstat=10)
        deck = os.path.join(td, "in.miniMD")
        open(deck, "w").write(cfg.input_text())
        cmd = [ref_binary("f64"), "-i", deck, "-f", data, "--half_neigh", str(half), "-gn", str(gn), "-o", "1", "--yaml_screen"]
        r = subprocess.run(cmd, cwd=REF_DIR, capture_output=True, text=True)
        for fn in os.listdir(REF_DIR):
            if fn.startswith("miniMD-") and fn.endswith(".yaml"):
                os.remove(os.path.join(REF_DIR, fn))
    res = parse_reference_output(r.stdout, yaml=True)
    assert res.steps, r.stdout[-2000:]
    import re
    bins = re.search(r"neighbor_bins: (\d+) (\d+) (\d+)", r.stdout)
    out[name] = {"half": half, "gn": gn, "steps": res.steps, "T": res.T, "U": res.U, "P": res.P, "nlocal": res.nlocal,
                 "nghost": res.nghost, "neighs": res.neighs, "bins": [int(b) for b in bins.groups()]}
    print(name, res.steps[-1], res.T[-1], res.U[-1], res.P[-1], res.nghost, res.neighs, out[name]["bins"])
json.dump(out, open(os.path.join(HERE, "reference_datafile.json"), "w"), indent=1)
