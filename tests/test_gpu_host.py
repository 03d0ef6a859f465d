"""GPU tests of the full product path: input deck -> C++ host layer -> C ABI -> CUDA kernels,
against the reference binary's golden runs (tests/golden/reference_runs.json: T/U/P with 10
significant digits, ghost and neighbor counts) and the reference's shipped logs -- including
BASELINE.json's full-size configurations.  FP64 tolerance on T/U/P: 1e-9 relative over 100 steps at
small sizes (summation order only), 1e-7 at the full sizes whose goldens were produced with 8
OpenMP threads; the reference's own acceptance is 1e-5."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import eam_file, g6, golden, rel
from minimd_b200 import Simulation, input_file
from minimd_b200.build import BIN
from oracle.oracle import Config, parse_reference_output

pytestmark = pytest.mark.gpu

RUNS = golden("reference_runs.json")
LOGS = golden("reference_logs.json")


def deck(tmp_path, cfg: Config):
    p = tmp_path / "in.miniMD"
    p.write_text(cfg.input_text())
    return str(p)


def args_for(tmp_path, cfg: Config, extra=()):
    c = cfg.resolved()
    a = ["-i", deck(tmp_path, c), "--half_neigh", c.halfneigh, "-gn", c.ghost_newton, "--sort", c.sort, "--ntypes", c.ntypes]
    if c.force == "eam":
        a += ["--eam_file", eam_file(str(tmp_path))]
    return a + ["--quiet"] + list(extra)


def check_thermo(sim, g, tol):
    st, T, U, P = sim.thermo()
    assert st == g["steps"]
    assert rel(T, g["T"]) < tol and rel(U, g["U"]) < tol
    assert np.max(np.abs(np.array(P) - np.array(g["P"]))) < 10 * tol * max(1.0, np.max(np.abs(g["P"])))


SMALL = ["lj_s8_half1_gn1", "lj_s8_half1_gn0", "lj_s8_half0_gn0", "eam_s8_half1_gn0", "eam_s8_half0_gn0",
         "lj_s8_every_step", "lj_6x8x10_nosort"]


@pytest.mark.parametrize("stepwise", [0, 1])
@pytest.mark.parametrize("name", SMALL)
def test_simulation_matches_reference_binary(name, stepwise, tmp_path):
    g = RUNS[name]
    cfg = Config(**g["config"])
    sim = Simulation(args_for(tmp_path, cfg, ["--stepwise"] if stepwise else []), g["precision"])
    sim.run()
    check_thermo(sim, g, 1e-9)
    assert sim.geti("nlocal") == g["nlocal"] and sim.geti("nghost") == g["nghost"]
    assert sim.geti("total_neigh") == g["neighs"]
    sim.finish()


def test_simulation_fp32_against_fp64_golden(tmp_path):
    g = RUNS["lj_s8_half0_gn0"]                      # FP64 golden; config 3 of BASELINE.json is full list, FP32
    sim = Simulation(args_for(tmp_path, Config(**g["config"])), "f32")
    sim.run()
    check_thermo(sim, g, 1e-4)


def test_simulation_long_run_follows_shipped_log(tmp_path):
    g = LOGS["4k.lj"]                                # 16-rank reference log, %e output (7 digits)
    nx, ny, nz = g["cells"]
    sim = Simulation(args_for(tmp_path, Config(nx=nx, ny=ny, nz=nz, ntimes=1000)), "f64")
    sim.run()
    st, T, U, P = sim.thermo()
    k = len(st)
    assert st == g["steps"][:k] and k == 11
    assert rel(T, g["T"][:k]) < 2e-6 and rel(U, g["U"][:k]) < 2e-6
    assert np.max(np.abs(np.array(P) - np.array(g["P"][:k]))) < 2e-6 * max(1.0, np.max(np.abs(g["P"][:k])))


@pytest.mark.parametrize("name", ["lj_s80_half", "lj_s80_full_gn0", "eam_s64_full", "lj_s32"])
def test_baseline_configs_full_size(name, tmp_path):
    """BASELINE.json configs 2-4 at full size: step-0 and step-100 T/U/P and the integer pins."""
    g = RUNS[name]
    sim = Simulation(args_for(tmp_path, Config(**g["config"])), g["precision"])
    sim.run()
    check_thermo(sim, g, 1e-7)
    # the reference prints its counts with %g: 6 significant digits
    assert g6(sim.geti("nghost")) == g["nghost"]
    assert g6(sim.geti("total_neigh")) == g["neighs"]
    assert sim.geti("natoms") == g["natoms"]


def test_config3_full_fp32_full_size(tmp_path):
    g = RUNS["lj_s80_full_gn0"]
    sim = Simulation(args_for(tmp_path, Config(**g["config"])), "f32")
    sim.run()
    check_thermo(sim, g, 1e-4)


def test_size_independent_invariants_at_full_size(tmp_path):
    """-s 80: momentum stays zero, every atom stays in the box, energy is conserved, and the half list is
    exactly half the full list (each pair stored once)."""
    cfg = Config(nx=80, ny=80, nz=80, ntimes=40, thermo_nstat=20)
    sim = Simulation(args_for(tmp_path, cfg), "f64")
    sim.run()
    c = sim.context()
    d = c.download("xv", count=sim.geti("nlocal"))
    assert np.abs(d["v"].sum(axis=0)).max() < 1e-6
    prd = np.array([sim.getr("xprd"), sim.getr("yprd"), sim.getr("zprd")])
    assert d["x"].min() > -1.0 and np.all(d["x"].max(axis=0) < prd + 1.0)
    st, T, U, P = sim.thermo()
    etot = [1.5 * t + u for t, u in zip(T, U)]
    # the first 40 steps off the perfect lattice are the roughest of the run (dt = 0.005): 3.4e-4 observed, same as the reference
    assert abs(etot[-1] - etot[0]) < 1e-3 * abs(etot[0])
    half_total = sim.geti("total_neigh")
    full = Simulation(args_for(tmp_path, Config(nx=80, ny=80, nz=80, ntimes=40, thermo_nstat=20, halfneigh=0, ghost_newton=0)), "f64")
    full.run()
    # same trajectory (to rounding) => the full list holds every half-list pair twice; allow a few pairs at the cutoff
    assert abs(full.geti("total_neigh") - 2 * half_total) <= 64
    _, T2, U2, P2 = full.thermo()
    assert rel(T2, T) < 1e-9 and rel(U2, U) < 1e-9


# --------------------------------------------------------------------------------------------------
# the drop-in driver executable: same command line, same screen output
# --------------------------------------------------------------------------------------------------
def run_driver(precision, args, cwd):
    exe = os.path.join(BIN, f"miniMD_b200_{precision}")
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    return subprocess.run([exe] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, timeout=600, env=env)


def test_driver_screen_output_is_reference_compatible(tmp_path):
    g = RUNS["lj_s8_half1_gn1"]
    cfg = Config(**g["config"])
    r = run_driver("f64", ["-i", deck(tmp_path, cfg), "-t", 1, "-dm", "--yaml_output", 0], str(tmp_path))
    assert r.returncode == 0, r.stderr
    out = r.stdout
    # tokens the reference's test script greps (ref/run_one_test:54-101)
    assert "# Timestep T U P Time" in out and "# Performance Summary" in out and "PERF_SUMMARY" in out
    m = re.search(r"# Size of float: (\d)", out)
    assert m and m.group(1) == "8"
    assert re.search(r"# Atoms: 2048\b", out) and "(unit cells: 8 8 8)" in out
    res = parse_reference_output(out, yaml=False)       # the parser used on the reference binary's own output
    assert res.steps == g["steps"]
    assert rel(res.T, g["T"]) < 1e-6 and rel(res.U, g["U"]) < 1e-6      # %e prints 7 digits
    assert res.natoms == 2048 and res.perf > 0


def test_driver_yaml_report_counts(tmp_path):
    g = RUNS["eam_s8_half0_gn0"]
    cfg = Config(**g["config"])
    r = run_driver("f64", ["-i", deck(tmp_path, cfg), "--half_neigh", 0, "--eam_file", eam_file(str(tmp_path)), "-o", 1,
                           "--yaml_screen"], str(tmp_path))
    assert r.returncode == 0, r.stderr
    res = parse_reference_output(r.stdout, yaml=True)
    assert res.steps == g["steps"]
    assert rel(res.T, g["T"]) < 2e-9 and rel(res.U, g["U"]) < 2e-9
    assert (res.nlocal, res.nghost, res.neighs) == (g["nlocal"], g["nghost"], g["neighs"])
    assert any(f.startswith("miniMD-") and f.endswith(".yaml") for f in os.listdir(tmp_path))


def test_driver_fp32_and_help(tmp_path):
    r = run_driver("f32", ["-i", input_file("in.lj.miniMD"), "-s", 6, "-n", 40, "--half_neigh", 0], str(tmp_path))
    assert r.returncode == 0 and "# Size of float: 4" in r.stdout
    h = run_driver("f64", ["-h"], str(tmp_path))
    assert h.returncode == 0 and "--half_neigh" in h.stdout
    bad = run_driver("f64", ["-i", "no_such_file"], str(tmp_path))
    assert bad.returncode != 0 and "Cannot open" in (bad.stdout + bad.stderr)


@pytest.mark.parametrize("name", ["lj_data_half", "lj_data_full"])
def test_lammps_data_file_run_matches_reference_binary(name, tmp_path):
    """Start from a LAMMPS data file (the reference's alternative input path, ref/setup.cpp:215-301)."""
    from helpers import lattice_for_datafile, write_lammps_data
    g = golden("reference_datafile.json")[name]
    x, v, prd = lattice_for_datafile()
    data = write_lammps_data(str(tmp_path / "atoms.data"), x, v, prd)
    cfg = Config(nx=6, ny=6, nz=6, ntimes=100, halfneigh=g["half"], ghost_newton=g["gn"], thermo_nstat=10)
    sim = Simulation(args_for(tmp_path, cfg, ["-f", data]), "f64")
    sim.run()
    check_thermo(sim, g, 1e-9)
    assert (sim.geti("nlocal"), sim.geti("nghost"), sim.geti("total_neigh")) == (g["nlocal"], g["nghost"], g["neighs"])
