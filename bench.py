#!/usr/bin/env python
"""Benchmark of the miniMD hot path on B200 (BASELINE.json: Matom-steps/s, in.lj.miniMD -s 80).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One bench "step" = one neighbor cycle of the reference's time loop = `neigh_every` (20) MD
timesteps: 19 steps of {initialIntegrate, forward halo, force, reverse halo, finalIntegrate} and one
re-neighboring step {exchange, sort, borders, neighbor build} -- so every timed step contains the
whole hot path in the proportions the reference runs it (ref/integrate.cpp:88-205).
Workload at N=1: BASELINE.json configs[1], `in.lj.miniMD -s 80` (2 048 000 atoms), half neighbor
list + ghost_newton, FP64.  N>1: weak scaling, 80^3 cells per GPU.

value   : natoms * MD steps / device time (CUDA events on the context's stream, max over ranks),
          state resident in HBM.
e2e     : the same neighbor cycle through the C ABI with HOST buffers: every step uploads x, v, type
          from pinned host memory (mmd_atom_upload), rebuilds ghosts + lists + forces as the
          reference's main() does after setup (ref/ljs.cpp:445-459), runs the 20 MD steps, and
          downloads x, v and the thermo sums; wall clock with synchronisation on both sides.
roofline: the force kernel (the dominant launch) -- algorithmic bytes per launch (SURVEY.md 8d row
          formulas with the neighbor count measured in this run) / its CUDA-event duration inside the
          timed region, against MEASURED_PEAKS.json's HBM copy bandwidth.
cpu_baseline / --impl reference: the UNMODIFIED reference binary (oracle/_ref, built by
          oracle/build_ref.sh) on this box's host cores, OpenMP over all of them, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MD_STEPS_PER_STEP = 20          # = neigh_every of in.lj.miniMD
CELLS_PER_GPU = 80
WEAK_BOX = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}
METRIC = "Matom-steps/s (device-timed) LJ -s 80 at 1/2/4/8 B200"
UNIT = "Matom-steps/s"


def env_int(k, d):
    v = os.environ.get(k)
    return int(v) if v not in (None, "") else d


def box_cells(n_gpus: int, size: int):
    m = WEAK_BOX.get(n_gpus)
    if m is None:                      # any other count: stack along z
        m = (1, 1, n_gpus)
    return size * m[0], size * m[1], size * m[2]


def workload_config(a, n_gpus):
    nx, ny, nz = box_cells(n_gpus, a.size)
    return {
        "workload": f"in.{a.force}.miniMD -s {a.size} per GPU ({nx}x{ny}x{nz} cells, {4 * nx * ny * nz} atoms), "
                    f"{'half' if a.half_neigh else 'full'} neighbor list, ghost_newton {a.ghost_newton if a.half_neigh else 0}, "
                    f"{'FP64' if a.precision == 'f64' else 'FP32'}",
        "natoms": 4 * nx * ny * nz, "cells": [nx, ny, nz], "half_neigh": a.half_neigh,
        "ghost_newton": a.ghost_newton if a.half_neigh else 0, "precision": a.precision,
        "md_steps_per_step": MD_STEPS_PER_STEP, "neigh_every": 20, "sort_every": 20, "thermo_every": 100,
        "decomposition": "x".join(str(v) for v in WEAK_BOX.get(n_gpus, (1, 1, n_gpus))),
        "l2": "working set (x,v,f + neighbor rows, ~0.6 GB per GPU) is several times the 126 MB L2; no flush between steps",
    }


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.rows, self.thread = index, None, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
            return
        def pump():
            for ln in self.proc.stdout:
                self.rows.append(ln.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            p = [t.strip() for t in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference binary on the host cores
# --------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_binary(cells, md_steps, half, gn, precision, threads, timeout=1500):
    """-> dict(value Matom-steps/s, t_total, kind, ...) using oracle/_ref (kind 'reference') or, if that
    binary is absent, the plain-C oracle port on one core (kind 'port')."""
    from oracle import oracle as orc
    cfg = orc.Config(nx=cells[0], ny=cells[1], nz=cells[2], ntimes=md_steps, halfneigh=half, ghost_newton=gn,
                     thermo_nstat=100)
    natoms = 4 * cells[0] * cells[1] * cells[2]
    if orc.ref_binary(precision):
        t0 = time.time()
        r = orc.run_reference(cfg, precision, threads=threads, yaml=False, timeout=timeout)
        if r.natoms > 0 and r.t_total > 0:
            return {"value": r.perf / 1e6, "t_total_s": r.t_total, "t_force_s": r.t_force, "t_neigh_s": r.t_neigh,
                    "t_comm_s": r.t_comm, "wall_s": time.time() - t0, "kind": "reference", "cores": threads,
                    "natoms": r.natoms, "md_steps": md_steps}
    o = orc.Oracle(cfg, precision)
    t0 = time.time()
    o.run(md_steps)
    dt = time.time() - t0
    return {"value": natoms * md_steps / dt / 1e6, "t_total_s": dt, "wall_s": dt, "kind": "port", "cores": 1,
            "natoms": natoms, "md_steps": md_steps}


def cpu_baseline(a, n_gpus, md_steps):
    """Half list (OpenMP atomics) and full list, all host cores; the faster one is 'ref CPU' (BASELINE.md section 2)."""
    cells = box_cells(n_gpus, a.size)
    threads = host_cores()
    runs = {}
    for name, half, gn in (("half", 1, 1), ("full", 0, 0)):
        try:
            runs[name] = run_reference_binary(cells, md_steps, half, gn, a.precision, threads)
        except Exception as e:  # noqa: BLE001
            runs[name] = {"value": 0.0, "error": repr(e), "kind": "reference", "cores": threads}
        if runs[name].get("kind") == "port":
            break               # the single-core port is slow: one style is enough
    best = max(runs, key=lambda k: runs[k]["value"])
    b = runs[best]
    return {"value": b["value"], "unit": UNIT, "cores": b.get("cores", threads), "kind": b.get("kind", "reference"),
            "sample": f"{cells[0]}x{cells[1]}x{cells[2]} cells ({4 * cells[0] * cells[1] * cells[2]} atoms), {md_steps} MD steps "
                      f"(one neighbor cycle = {MD_STEPS_PER_STEP}), ref/ built by oracle/build_ref.sh, OpenMP threads = cores, "
                      f"timed by its own PERF_SUMMARY; best of half/full = {best}",
            "half_list": runs.get("half"), "full_list": runs.get("full")}


def reference_arm(a, n_gpus, rank):
    if rank != 0:
        return
    blocks = max(1, min(a.steps, 2))
    md_steps = MD_STEPS_PER_STEP * blocks
    cb = cpu_baseline(a, n_gpus, md_steps)
    best = cb["half_list"] if cb["half_list"] and cb["half_list"]["value"] >= cb["value"] else cb["full_list"]
    t_total = (best or {}).get("t_total_s", float("nan"))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * t_total / blocks, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.precision, "data": "synthetic", "config": workload_config(a, n_gpus),
            "cpu_baseline": cb, "gpu_launches": 0,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": f"reference's own CPU path: {blocks} neighbor cycle(s) in one process (its untimed warm-up build+force "
                    f"is the warm-up); --steps/--warmup beyond that are not repeated to keep the run bounded"}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------------
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(key):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:  # noqa: BLE001
            return None
    return None


def own_arm(a, n_gpus, rank, local_rank):
    import numpy as np
    import torch

    from minimd_b200 import Simulation, input_file, nccl_unique_id
    from minimd_b200._lib import check

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if n_gpus > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())
    else:
        nccl_id = None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nx, ny, nz = box_cells(n_gpus, a.size)
    total_md = MD_STEPS_PER_STEP * (a.warmup + a.steps)
    args = ["-i", input_file("in.lj.miniMD" if a.force == "lj" else "in.eam.miniMD"), "-nx", nx, "-ny", ny, "-nz", nz, "-n", total_md,
            "--half_neigh", a.half_neigh, "-gn", a.ghost_newton, "--quiet"]
    if a.force == "eam":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import eam_file            # the Cu_u6 table: oracle/_ref copy or the committed fixture
        args += ["--eam_file", eam_file()]
        a.no_e2e = True                          # the e2e leg drives the LJ entry points
    sim = Simulation(args, a.precision, rank=rank, nranks=n_gpus, device=local_rank, nccl_id=nccl_id)
    ctx = sim.context()
    if a.tpa:
        ctx.set_option("lj_threads_per_atom", a.tpa)
    ctx.set_option("tile_lists", a.tile)      # effective from the next neighbor build (inside the warm-up)
    if not a.p2p:
        ctx.set_option("p2p_halo", 0)
    for kv in a.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    natoms = sim.geti("natoms")
    stream = torch.cuda.ExternalStream(ctx.stream)

    # ---- resident run: W warm-up cycles, then K timed cycles -----------------------------------
    for _ in range(a.warmup):
        sim.run(MD_STEPS_PER_STEP)
    ctx.phase_times(reset=True)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    inner_ms = 0.0
    for _ in range(a.steps):
        inner_ms += sim.run(MD_STEPS_PER_STEP)
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    inner_ms = max_over_ranks(inner_ms)
    launches = ctx.launches - launches0
    clk = clocks.stop() if rank == 0 else None
    phases = ctx.phase_times()
    md_steps = MD_STEPS_PER_STEP * a.steps
    value = natoms * md_steps / (ms_total * 1e-3) / 1e6

    # ---- roofline of the force kernel ------------------------------------------------------------
    s = 8 if a.precision == "f64" else 4
    nlocal = sim.geti("nlocal")
    n_per_atom = sim.geti("total_neigh") / max(nlocal, 1)
    tiled = bool(ctx.query("list_tile"))
    f_bytes = (3 * s + 6 * s) if a.half_neigh else 3 * s        # half: clear + read-modify-write; full: one store
    force_bytes = nlocal * ((4 * n_per_atom + 4) + (3 * s + 4) + f_bytes)
    # tile-resident lists: the force launch of every step but the last of an mmd_run call also performs that step's
    # finalIntegrate and the next step's initialIntegrate (15 s + 9 s bytes per atom)
    fused_verlet = tiled and a.force == "lj" and bool(ctx.query("fuse_force"))
    if fused_verlet:
        force_bytes += nlocal * 24 * s * (MD_STEPS_PER_STEP - 1) / MD_STEPS_PER_STEP
    f_ms, f_calls = phases["force"]
    peak, peak_src = measured_peak()
    force_avg_ms = f_ms / max(f_calls, 1)
    achieved = force_bytes / (force_avg_ms * 1e-3) / 1e9 if force_avg_ms > 0 else 0.0
    kern = (f"force_{a.force}_{'tile_' if tiled else ''}kernel<{'double' if s == 8 else 'float'},half={a.half_neigh},"
            f"gn={a.ghost_newton if a.half_neigh else 0}>")
    if a.force == "eam":                                          # two passes over the rows + fp traffic
        force_bytes = nlocal * (2 * ((4 * n_per_atom + 4) + (3 * s + 4)) + f_bytes + 2 * s)
    roofline = {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": profiled_traffic("force_lj_tile_f64" if tiled else ("force_lj_half_f64" if a.half_neigh else "force_lj_full_f64")),
                "list_format": "tile-resident 16-bit rows, owner-computes" if tiled else "classic rows of global ids",
                "fused_verlet": fused_verlet,
                "algorithmic_bytes_per_launch": force_bytes, "avg_launch_ms": force_avg_ms, "launches_timed": f_calls,
                "neighbors_per_atom": n_per_atom, "peak_source": peak_src,
                "share_of_step": f_ms / max(sum(v[0] for v in phases.values()), 1e-12)}
    # whole neighbor cycle against the same roof (sum of the per-kernel rows)
    rebuild = ((3 * s + 8) + (3 * s + 8 + 4 * n_per_atom) + 2 * (6 * s + 4)) / 20.0
    step_bytes = (4 * n_per_atom + 4) + (3 * s + 4) + f_bytes + 15 * s + 9 * s + rebuild
    per_gpu_rate = value * 1e6 / n_gpus
    step_roofline = {"bytes_per_atom_step": step_bytes, "achieved": per_gpu_rate * step_bytes / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": per_gpu_rate * step_bytes / 1e9 / peak}
    phase_ms = {k: v[0] / a.steps for k, v in phases.items()}

    # ---- end to end through the C ABI with host buffers ---------------------------------------------
    e2e = None
    if not a.no_e2e:
        real = np.float64 if s == 8 else np.float32
        cap = int(nlocal * 1.1) + 4096
        hx = torch.empty((cap, 3), dtype=torch.float64 if s == 8 else torch.float32).pin_memory().numpy()
        hv = torch.empty((cap, 3), dtype=torch.float64 if s == 8 else torch.float32).pin_memory().numpy()
        ht = torch.empty((cap,), dtype=torch.int32).pin_memory().numpy()
        lib = ctx.lib
        import ctypes as C
        vp = lambda arr: arr.ctypes.data_as(C.c_void_p)
        half, gn = a.half_neigh, (a.ghost_newton if a.half_neigh else 0)
        params = sim.run_params(MD_STEPS_PER_STEP)
        from minimd_b200._lib import ThermoSample
        samples = (ThermoSample * 4)()
        nsamp = C.c_int()

        def download(n):
            check(lib.mmd_atom_download(ctx.h, vp(hx), None, None, vp(ht), 0, n, 3))
            check(lib.mmd_atom_download(ctx.h, None, vp(hv), None, None, 0, n, 3))

        n_now = ctx.counts()[0]
        download(n_now)
        h2d = d2h = 0

        def cycle(n):
            # the reference's start-up sequence after setup (ref/ljs.cpp:445-459), then 20 steps
            check(lib.mmd_atom_upload(ctx.h, vp(hx), vp(hv), vp(ht), n, 3))
            check(lib.mmd_comm_exchange(ctx.h))
            check(lib.mmd_comm_borders(ctx.h))
            mx, tot = C.c_int(0), C.c_longlong()
            check(lib.mmd_neigh_build(ctx.h, half, gn, C.byref(mx), C.byref(tot)))
            check(lib.mmd_force_lj_compute(ctx.h, half, gn, 0, None, None))
            if half and gn:
                check(lib.mmd_comm_reverse_communicate(ctx.h))
            check(lib.mmd_run(ctx.h, C.byref(params), samples, 4, C.byref(nsamp), None))
            n2 = ctx.counts()[0]
            download(n2)
            return n2

        n_now = cycle(n_now)          # warm-up (allocations, staging buffers)
        barrier()
        t0 = time.perf_counter()
        reps = max(1, min(a.steps, 10))
        for _ in range(reps):
            h2d += n_now * (2 * 3 * s + 4)
            n_now = cycle(n_now)
            d2h += n_now * (2 * 3 * s + 4) + 4 * 32
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": natoms * MD_STEPS_PER_STEP * reps / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(sum_over_ranks(h2d) / reps), "d2h_bytes_per_step": int(sum_over_ranks(d2h) / reps),
               "steps": reps, "ms_per_step": 1e3 * dt / reps,
               "path": "mmd_atom_upload(host x,v,type) -> exchange -> borders -> neigh_build -> force -> reverse -> "
                       "mmd_run(20 MD steps) -> mmd_atom_download(host x,v,type)"}

    st, T, U, P = sim.thermo()
    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.precision, "data": "synthetic", "config": workload_config(a, n_gpus),
        "md_steps_timed": md_steps, "device_ms_inside_mmd_run": inner_ms, "phase_ms_per_step": phase_ms,
        "roofline": roofline, "step_roofline": step_roofline, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clk, "thermo_last": {"step": st[-1], "T": T[-1], "U": U[-1], "P": P[-1]} if st else None,
        "counts": {"nlocal_rank0": nlocal, "nghost_rank0": sim.geti("nghost"), "maxneighs": sim.geti("maxneighs")},
        "halo_transport": ("none (single rank: device-local self swaps)" if n_gpus == 1 else
                           ("peer-memory windows over NVLink (CUDA IPC), fused pack+remote store / wait+unpack kernels"
                            if ctx.query("p2p_active") else "NCCL send/recv + pack/unpack kernels")),
    }
    if rank == 0 and n_gpus == 1 and not a.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline(a, 1, MD_STEPS_PER_STEP)
    elif rank == 0:
        result["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(result), flush=True)
    sim.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--size", type=int, default=CELLS_PER_GPU, help="unit cells per GPU edge (the metric is quoted at 80)")
    ap.add_argument("--half_neigh", type=int, default=1)
    ap.add_argument("--ghost_newton", type=int, default=1)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--force", default="lj", choices=["lj", "eam"], help="lj = the headline metric; eam = BASELINE.json config 4 (use --size 64)")
    ap.add_argument("--tpa", type=int, default=0, help="lanes per atom in the force kernel (0 = library default)")
    ap.add_argument("--tile", type=int, default=1, help="1: tile-resident neighbor lists + shared-memory force kernel (default); "
                                                        "0: classic rows of global ids (gather / scatter kernels)")
    ap.add_argument("--p2p", type=int, default=1, help="N>1: 1 = forward halo over peer-memory windows (default), 0 = NCCL send/recv")
    ap.add_argument("--opt", action="append", default=[], help="library switch key=value (mmd_set_option), repeatable; for A/B runs")
    ap.add_argument("--no-e2e", dest="no_e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "own" else a.warmup
    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    n_gpus = a.gpus if world == 1 else world
    if world == 1 and a.gpus > 1:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if a.impl == "reference":
        reference_arm(a, n_gpus, rank)
        return
    own_arm(a, n_gpus, rank, local_rank)


if __name__ == "__main__":
    main()
