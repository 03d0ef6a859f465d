"""Pins the oracle (oracle/minimd_oracle.c) to the reference's own golden vectors:
  * the logs the reference ships under tests/reference_output (committed as
    tests/golden/reference_logs.json by tests/golden/make_golden.py) -- 16 MPI ranks, FP64;
  * the unmodified reference binary run in the build container on BASELINE.json's configs
    (tests/golden/reference_runs.json, 10 significant digits from its YAML report);
  * when oracle/_ref/ exists, a live run of that binary with thermo output at every step.
"""
import numpy as np
import pytest

from helpers import g6, golden, rel
from oracle.oracle import Config, Oracle, ref_binary, run_reference

LOGS = golden("reference_logs.json")
RUNS = golden("reference_runs.json")


def run_oracle(cfg, precision="f64"):
    o = Oracle(cfg, precision)
    o.run(cfg.ntimes)
    return o, o.thermo_log()


# the shipped logs print %e (7 significant digits) => 5e-7 relative is "all digits"
@pytest.mark.parametrize("log,nsteps", [("4k.lj", 1000), ("16k.lj", 200), ("32k.lj", 100), ("4k.eam", 300), ("16k.eam", 100)])
def test_oracle_matches_shipped_reference_logs(log, nsteps):
    g = LOGS[log]
    nx, ny, nz = g["cells"]
    force = "eam" if "eam" in log else "lj"
    o, (st, T, U, P) = run_oracle(Config(nx=nx, ny=ny, nz=nz, ntimes=nsteps, force=force, thermo_nstat=100))
    assert o.geti("natoms") == g["natoms"]
    k = len(st)
    assert list(st) == g["steps"][:k]
    assert rel(T, g["T"][:k]) < 1e-6
    assert rel(U, g["U"][:k]) < 1e-6
    # P crosses zero; compare on the scale of the kinetic pressure
    assert np.max(np.abs(np.array(P) - np.array(g["P"][:k]))) < 1e-6 * max(1.0, np.max(np.abs(g["P"][:k])))


SMALL = ["lj_s8_half1_gn1", "lj_s8_half1_gn0", "lj_s8_half0_gn0", "eam_s8_half1_gn1", "eam_s8_half1_gn0",
         "eam_s8_half0_gn0", "lj_s8_every_step", "lj_6x8x10_nosort", "lj_s8_f32_full"]


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_reference_binary_fixture(name):
    g = RUNS[name]
    cfg = Config(**g["config"])
    o, (st, T, U, P) = run_oracle(cfg, g["precision"])
    assert list(st) == g["steps"]
    # the YAML report prints 10 significant digits
    assert rel(T, g["T"]) < 2e-9 and rel(U, g["U"]) < 2e-9
    assert np.max(np.abs(np.array(P) - np.array(g["P"]))) < 2e-9 * max(1.0, np.max(np.abs(g["P"])))
    assert o.nlocal == g["nlocal"] and o.nghost == g["nghost"]
    assert int(o.numneigh().sum()) == g["neighs"]


def test_oracle_matches_reference_fixture_s32():
    g = RUNS["lj_s32"]
    o, (st, T, U, P) = run_oracle(Config(**g["config"]))
    # the fixture ran with 8 OpenMP threads (atomics => different summation order): 1e-9, not bitwise
    assert rel(T, g["T"]) < 1e-8 and rel(U, g["U"]) < 1e-8 and rel(P, g["P"]) < 1e-6
    # the reference prints its counts with %g (ref/output.cpp:402,458): 6 significant digits
    assert g6(o.nghost) == g["nghost"] and g6(int(o.numneigh().sum())) == g["neighs"]


@pytest.mark.skipif(ref_binary() is None, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("force,half,gn,prec", [("lj", 1, 1, "f64"), ("lj", 0, 0, "f64"), ("eam", 1, 0, "f64"),
                                                ("eam", 0, 0, "f64"), ("lj", 1, 1, "f32"), ("eam", 0, 0, "f32")])
def test_oracle_matches_live_reference_every_step(force, half, gn, prec):
    cfg = Config(nx=6, ny=6, nz=6, ntimes=45, force=force, halfneigh=half, ghost_newton=gn, thermo_nstat=1)
    ref = run_reference(cfg, prec)
    o, (st, T, U, P) = run_oracle(cfg, prec)
    assert list(st) == ref.steps
    assert rel(T, ref.T) < 2e-9 and rel(U, ref.U) < 2e-9
    assert np.max(np.abs(np.array(P) - np.array(ref.P))) < 2e-9 * max(1.0, np.max(np.abs(ref.P)))
    assert o.nghost == ref.nghost and int(o.numneigh().sum()) == ref.neighs
