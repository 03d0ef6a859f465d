"""Drop-in proof: the REFERENCE program -- its own main(), input, setup, Comm, Thermo, output --
with one translation unit replaced by a file that calls the C ABI (tests/dropin/*.cpp, built by oracle/build_dropin.sh into
oracle/_ref/).  The binaries print the reference's own thermo lines / YAML report; they are compared with the oracle.
Skipped where the binaries do not exist (they are built in the container that has /root/reference and travel with gpurun)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.oracle import Config, Oracle

pytestmark = pytest.mark.gpu


def run_dropin(exe_name, cfg, threads=1):
    exe = os.path.join(orc.REF_DIR, exe_name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe_name} not built (oracle/build_dropin.sh needs the reference tree)")
    c = cfg.resolved()
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "in.miniMD")
        with open(inp, "w") as fh:
            fh.write(c.input_text())
        cmd = [exe, "-i", inp, "-t", str(threads), "--half_neigh", str(c.halfneigh), "-gn", str(c.ghost_newton),
               "--sort", str(c.sort), "--ntypes", str(c.ntypes), "-o", "1", "--yaml_screen"]
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        r = subprocess.run(cmd, cwd=orc.REF_DIR, capture_output=True, text=True, timeout=600, env=env)
        for fn in os.listdir(orc.REF_DIR):
            if fn.startswith("miniMD-") and fn.endswith(".yaml"):
                os.remove(os.path.join(orc.REF_DIR, fn))
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "# Performance Summary" in r.stdout and "PERF_SUMMARY" in r.stdout
    return orc.parse_reference_output(r.stdout, yaml=True)


def check_against_oracle(res, cfg, tol):
    o = Oracle(cfg, "f64")
    o.run(cfg.ntimes)
    st, T, U, P = o.thermo_log()
    assert list(res.steps) == list(st), (res.steps, list(st))
    pscale = max(1.0, float(np.max(np.abs(P))))
    for k in range(len(st)):
        assert abs(res.T[k] - T[k]) <= tol * abs(T[k]), (st[k], res.T[k], T[k])
        assert abs(res.U[k] - U[k]) <= tol * abs(U[k]), (st[k], res.U[k], U[k])
        assert abs(res.P[k] - P[k]) <= 10 * tol * pscale, (st[k], res.P[k], P[k])
    assert res.natoms == o.geti("natoms")


@pytest.mark.parametrize("force,half,gn,size", [("lj", 1, 1, 8), ("lj", 0, 0, 6), ("eam", 0, 0, 6), ("eam", 1, 0, 6)])
def test_reference_main_with_integrate_run_on_the_device(force, half, gn, size):
    """ref/integrate.cpp replaced: the reference's main() drives mmd_run; its own Thermo prints the lines."""
    cfg = Config(nx=size, ny=size, nz=size, ntimes=100, force=force, halfneigh=half, ghost_newton=gn, thermo_nstat=20)
    res = run_dropin("miniMD_dropin_run_f64", cfg)
    check_against_oracle(res, cfg, 2e-9)


@pytest.mark.parametrize("half,gn", [(1, 1), (1, 0), (0, 0)])
def test_reference_time_loop_with_force_compute_on_the_device(half, gn):
    """ref/force_lj.cpp replaced: the reference's own time loop, Comm and Neighbor on the host; every ForceLJ::compute is
    upload -> mmd_force_lj_compute on the host-built list -> download (the mpi-spec compute_lj seam)."""
    cfg = Config(nx=6, ny=6, nz=6, ntimes=60, halfneigh=half, ghost_newton=gn, thermo_nstat=20)
    res = run_dropin("miniMD_dropin_force_f64", cfg)
    check_against_oracle(res, cfg, 2e-9)
