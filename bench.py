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
roofline: the force launch of one MD step (the dominant kernel; with tile lists it also performs the
          two velocity-Verlet halves) -- algorithmic bytes per launch (SURVEY.md 8d row formulas with
          the neighbor count measured in this run) / its CUDA-event duration inside the timed region,
          against MEASURED_PEAKS.json's HBM copy bandwidth.  (One rank replays pairs of plain steps from
          a CUDA graph: the events then sit around a pair, and the force launch is charged with the two
          7-us halo launches inside -- `halo_in_force_interval`.)  `fp64_frac`: pair evaluations per second
          against the measured FP64 pair rate (profiles/r2_fp64_peak.json, tools/microbench/
          fp64_fma_bench.cu).  `traffic`: DRAM bytes per launch from the committed ncu capture of THIS
          kernel on THIS workload (profiles/traffic.json names kernel, atoms and source) or null.
other_configs (N=1): BASELINE.json configs[2] (`-s 80` full list FP32) and configs[3] (EAM `-s 64`
          full list FP64), a few cycles each, each with its own roofline / e2e / cpu_baseline.
cpu_baseline / --impl reference: the UNMODIFIED reference binary (oracle/_ref, built by
          oracle/build_ref.sh) on this box's host cores, OpenMP over all of them, bounded sample;
          the reference arm runs --warmup + --steps cycles in one process (or fewer if they would
          not end within ~150 s, and then says so in `steps`).
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


def run_reference_binary(cells, md_steps, half, gn, precision, threads, force="lj", timeout=1500):
    """-> dict(value Matom-steps/s, t_total, kind, ...) using oracle/_ref (kind 'reference') or, if that
    binary is absent, the plain-C oracle port on one core (kind 'port')."""
    from oracle import oracle as orc
    cfg = orc.Config(nx=cells[0], ny=cells[1], nz=cells[2], ntimes=md_steps, force=force, halfneigh=half, ghost_newton=gn,
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


def cpu_baseline(a, n_gpus, md_steps, styles=None):
    """The reference on all host cores, bounded sample.  LJ: half list (OpenMP atomics) and full list, the faster one is
    'ref CPU' (BASELINE.md section 2); other configurations: the list style the configuration names."""
    cells = box_cells(n_gpus, a.size)
    threads = host_cores()
    if styles is None:
        styles = (("half", 1, 1), ("full", 0, 0))
    runs = {}
    for name, half, gn in styles:
        try:
            runs[name] = run_reference_binary(cells, md_steps, half, gn, a.precision, threads, a.force)
        except Exception as e:  # noqa: BLE001
            runs[name] = {"value": 0.0, "error": repr(e), "kind": "reference", "cores": threads}
        if runs[name].get("kind") == "port":
            break               # the single-core port is slow: one style is enough
    best = max(runs, key=lambda k: runs[k]["value"])
    b = runs[best]
    return {"value": b["value"], "unit": UNIT, "cores": b.get("cores", threads), "kind": b.get("kind", "reference"),
            "sample": f"in.{a.force}.miniMD {cells[0]}x{cells[1]}x{cells[2]} cells ({4 * cells[0] * cells[1] * cells[2]} atoms), "
                      f"{md_steps} MD steps (one neighbor cycle = {MD_STEPS_PER_STEP}), {'FP64' if a.precision == 'f64' else 'FP32'}, "
                      f"ref/ built by oracle/build_ref.sh, OpenMP threads = cores, timed by its own PERF_SUMMARY; "
                      f"list style = {best}" + (" (best of half/full)" if len(runs) > 1 else ""),
            "best_style": best, "half_list": runs.get("half"), "full_list": runs.get("full")}


def reference_arm(a, n_gpus, rank):
    """The reference's own CPU implementation, `--warmup + --steps` neighbor cycles in ONE process of the unmodified binary
    (its PERF_SUMMARY times the whole loop, so the rate is the average over all W+K cycles; ms_per_step = loop time /
    (W+K)).  A one-cycle probe of each list style picks the faster one and sizes the run: if W+K cycles would not end
    within ~150 s, fewer are run and `steps` reports what really ran."""
    if rank != 0:
        return
    probe = cpu_baseline(a, n_gpus, MD_STEPS_PER_STEP)
    style = probe["best_style"]
    half, gn = (1, 1) if style == "half" else (0, 0)
    cyc_s = max((probe["half_list" if style == "half" else "full_list"] or {}).get("t_total_s", 1.0), 1e-3)
    want = a.steps + a.warmup
    cycles = max(1, min(want, int(150.0 / cyc_s)))
    steps_run = a.steps if cycles == want else max(1, cycles - min(a.warmup, cycles - 1))
    warm_run = cycles - steps_run
    cells = box_cells(n_gpus, a.size)
    r = run_reference_binary(cells, MD_STEPS_PER_STEP * cycles, half, gn, a.precision, host_cores(), a.force)
    cb = {"value": r["value"], "unit": UNIT, "cores": r.get("cores", 1), "kind": r.get("kind", "reference"),
          "sample": f"{cells[0]}x{cells[1]}x{cells[2]} cells ({r['natoms']} atoms), {cycles} neighbor cycles = {r['md_steps']} MD steps "
                    f"in one process, {style} neighbor list (the faster of half/full in a one-cycle probe: "
                    f"half {((probe['half_list'] or {}).get('value') or 0):.2f}, full {((probe['full_list'] or {}).get('value') or 0):.2f} "
                    f"Matom-steps/s), OpenMP threads = cores, timed by its own PERF_SUMMARY",
          "run": r, "probe": probe}
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": steps_run,
            "warmup": warm_run, "ms_per_step": 1e3 * r["t_total_s"] / cycles, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.precision, "data": "synthetic", "config": workload_config(a, n_gpus),
            "cpu_baseline": cb, "gpu_launches": 0,
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": f"reference's own CPU path (oracle/_ref, unmodified): {cycles} neighbor cycles timed as one loop by the "
                    f"binary itself; requested --steps {a.steps} --warmup {a.warmup}"}
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


def fp64_peak():
    """Measured vector-FP64 roof (tools/microbench/fp64_fma_bench.cu on a B200 of this pool, kept under profiles/)."""
    p = os.path.join(ROOT, "profiles", "r2_fp64_peak.json")
    try:
        return json.load(open(p))
    except Exception:  # noqa: BLE001
        return None


def profiled_traffic(kernel_key, natoms):
    """DRAM bytes per launch of the named kernel from the committed `ncu --set full` capture -- only when that capture
    is of THIS kernel on THIS workload (profiles/traffic.json records kernel, atoms and the library build it was taken
    with); anything else reads null rather than a stale constant."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        e = json.load(open(p)).get(kernel_key)
    except Exception:  # noqa: BLE001
        return None, None
    if not isinstance(e, dict) or e.get("natoms") != natoms:
        return None, None
    return e.get("dram_bytes_per_launch"), e.get("source")


def force_kernel_bytes(force, s, half, n_per_atom, fused_verlet_share):
    """Algorithmic bytes per atom of one force launch."""
    f_bytes = (3 * s + 6 * s) if half else 3 * s            # half: clear + read-modify-write; full: one store
    if force == "eam":                                        # two passes over the rows + fp write/read
        return 2 * ((4 * n_per_atom + 4) + (3 * s + 4)) + f_bytes + 2 * s
    b = (4 * n_per_atom + 4) + (3 * s + 4) + f_bytes
    # tile-resident lists: the launch also performs finalIntegrate(n) + initialIntegrate(n+1) (rows a13/a14: 15 s + 9 s)
    return b + 24 * s * fused_verlet_share


class Dist:
    """torch.distributed plumbing of the bench (rendezvous, barrier, reductions); the data path never touches it."""

    def __init__(self, n_gpus, rank, local_rank):
        import torch
        self.torch = torch
        self.n, self.rank = n_gpus, rank
        self.dist = None
        self.nccl_id = None
        if n_gpus > 1:
            import torch.distributed as dist_
            from minimd_b200 import nccl_unique_id
            self.dist = dist_
            dist_.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt = torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8).cuda()
            dist_.broadcast(idt, 0)
            self.nccl_id = bytes(idt.cpu().numpy().tobytes())

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _red(self, v, op):
        if self.dist is None:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v):
        return self._red(v, self.dist.ReduceOp.MAX) if self.dist else v

    def sum(self, v):
        return self._red(v, self.dist.ReduceOp.SUM) if self.dist else v

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def make_sim(a, D, local_rank):
    from minimd_b200 import Simulation, input_file
    nx, ny, nz = box_cells(D.n, a.size)
    total_md = MD_STEPS_PER_STEP * (a.warmup + a.steps)
    args = ["-i", input_file("in.lj.miniMD" if a.force == "lj" else "in.eam.miniMD"), "-nx", nx, "-ny", ny, "-nz", nz, "-n", total_md,
            "--half_neigh", a.half_neigh, "-gn", a.ghost_newton, "--quiet"]
    if a.force == "eam":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import eam_file            # the Cu_u6 table: oracle/_ref copy or the committed fixture
        args += ["--eam_file", eam_file()]
    sim = Simulation(args, a.precision, rank=D.rank, nranks=D.n, device=local_rank, nccl_id=D.nccl_id)
    ctx = sim.context()
    if a.tpa:
        ctx.set_option("lj_threads_per_atom", a.tpa)
    ctx.set_option("tile_lists", a.tile)      # effective from the next neighbor build (inside the warm-up)
    if not a.p2p:
        ctx.set_option("p2p_halo", 0)
    for kv in a.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    return sim, ctx


def measure_resident(a, D, sim, ctx, local_rank, with_clocks):
    """W warm-up cycles, then K timed cycles with the state resident in HBM; CUDA events on the context's stream."""
    torch = D.torch
    natoms = sim.geti("natoms")
    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(a.warmup):
        sim.run(MD_STEPS_PER_STEP)
    ctx.phase_times(reset=True)
    clocks = ClockSampler(local_rank)
    if with_clocks:
        clocks.start()
    D.barrier()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    inner_ms = 0.0
    for _ in range(a.steps):
        inner_ms += sim.run(MD_STEPS_PER_STEP)
    e1.record(stream)
    e1.synchronize()
    D.barrier()
    ms_total = D.max(e0.elapsed_time(e1))
    inner_ms = D.max(inner_ms)
    launches = ctx.launches - launches0
    clk = clocks.stop() if with_clocks else None
    phases = ctx.phase_times()
    md_steps = MD_STEPS_PER_STEP * a.steps
    value = natoms * md_steps / (ms_total * 1e-3) / 1e6

    # ---- roofline of the force launch(es) of one MD step -------------------------------------------
    s = 8 if a.precision == "f64" else 4
    nlocal = sim.geti("nlocal")
    n_per_atom = sim.geti("total_neigh") / max(nlocal, 1)
    tiled = bool(ctx.query("list_tile"))
    dealt = bool(ctx.query("list_dealt"))
    fused_verlet = tiled and a.force == "lj" and bool(ctx.query("fuse_force"))
    share = (MD_STEPS_PER_STEP - 1) / MD_STEPS_PER_STEP if fused_verlet else 0.0
    per_atom = force_kernel_bytes(a.force, s, a.half_neigh, n_per_atom, share)
    force_bytes = nlocal * per_atom
    f_ms, f_calls = phases["force"]
    split_steps = ctx.query("split_steps")
    if split_steps > 0:
        # several ranks: interior tiles run on a second stream behind the forward halo, so the stream the phase events sit
        # on sees halo + wait + boundary tiles; force and comm are timed TOGETHER (an upper bound of the force launch)
        f_ms += phases["comm"][0]
    peak, peak_src = measured_peak()
    force_avg_ms = f_ms / max(f_calls, 1)
    achieved = force_bytes / (force_avg_ms * 1e-3) / 1e9 if force_avg_ms > 0 else 0.0
    real = "double" if s == 8 else "float"
    if a.force == "lj":
        kern = ("force_lj_dealt_kernel" if dealt else ("force_lj_tile_kernel" if tiled else "force_lj_kernel")) + \
               f"<{real},half={a.half_neigh},gn={a.ghost_newton if a.half_neigh else 0}>"
        tkey = f"force_lj_{'dealt' if dealt else ('tile' if tiled else 'classic')}_{a.precision}_{'half' if a.half_neigh else 'full'}"
        fmt = ("tile-resident 16-bit rows, bank-dealt, quarter warp per atom, owner-computes" if dealt else
               ("tile-resident 16-bit rows, owner-computes" if tiled else "classic rows of global ids"))
    else:
        kern = ("eam_dealt_kernel<1>+<2>" if dealt else ("eam_tile_kernel<1>+<2>" if tiled else "eam_rho+embed+pair_kernel")) + \
               f"<{real},half={a.half_neigh}> (+ fp halo)"
        tkey = f"force_eam_{'dealt' if dealt else ('tile' if tiled else 'classic')}_{a.precision}_{'half' if a.half_neigh else 'full'}"
        fmt = "tile-resident rows" if tiled else "classic rows of global ids"
    traffic, traffic_src = profiled_traffic(tkey, natoms)
    roofline = {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "list_format": fmt, "fused_verlet": fused_verlet,
                "halo_overlapped_in_timing": bool(split_steps > 0),
                # one rank: pairs of plain steps are replayed from a CUDA graph; the events sit around a pair, so the force
                # interval then also holds the two 7-us halo launches (the force launch is charged with them)
                "halo_in_force_interval": bool(ctx.query("graph_replays") > 0),
                "algorithmic_bytes_per_atom": per_atom, "algorithmic_bytes_per_launch": force_bytes,
                "avg_launch_ms": force_avg_ms, "launches_timed": f_calls,
                "neighbors_per_atom": n_per_atom, "peak_source": peak_src,
                "share_of_step": f_ms / max(sum(v[0] for v in phases.values()), 1e-12)}
    # second roof: the FP64 pipe.  Pair evaluations of one launch (owner-computes visits every stored pair from both
    # ends; the classic half-list kernels once) against the measured pair rate of tools/microbench/fp64_fma_bench.cu
    fp = fp64_peak()
    if fp and s == 8 and a.force == "lj" and force_avg_ms > 0:
        evals = nlocal * n_per_atom * ((2.0 if a.half_neigh else 1.0) if tiled else 1.0)
        roofline["fp64_pair_evals_per_launch"] = evals
        roofline["fp64_peak_pairs_per_s"] = fp.get("lj_pairs_per_s")
        roofline["fp64_frac"] = evals / (force_avg_ms * 1e-3) / fp["lj_pairs_per_s"] if fp.get("lj_pairs_per_s") else None
        roofline["fp64_peak_source"] = "profiles/r2_fp64_peak.json (measured, tools/microbench/fp64_fma_bench.cu)"
    # whole neighbor cycle against the same roof (sum of the per-kernel rows)
    rebuild = ((3 * s + 8) + (3 * s + 8 + 4 * n_per_atom) + 2 * (6 * s + 4)) / 20.0
    step_bytes = force_kernel_bytes(a.force, s, a.half_neigh, n_per_atom, 0.0) + 15 * s + 9 * s + rebuild
    per_gpu_rate = value * 1e6 / D.n
    step_roofline = {"bytes_per_atom_step": step_bytes, "achieved": per_gpu_rate * step_bytes / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": per_gpu_rate * step_bytes / 1e9 / peak}
    kp = None
    if any(kv.startswith("kernel_profile=") and kv != "kernel_profile=0" for kv in a.opt):
        n_cta = ctx.query("cta_count")
        if n_cta > 0:   # clock64 sums per CTA of the LJ tile force kernel (diagnostic run: not a bench number)
            kp = {"ctas": n_cta, "stage_clocks_per_cta": ctx.query("stage_clocks") / n_cta,
                  "cta_clocks": ctx.query("cta_clocks") / n_cta,
                  "staging_share_of_cta_lifetime": ctx.query("stage_clocks") / max(ctx.query("cta_clocks"), 1)}
    st, T, U, P = sim.thermo()
    return {
        "kernel_profile": kp,
        "value": value, "ms_per_step": ms_total / a.steps, "md_steps_timed": md_steps, "device_ms_inside_mmd_run": inner_ms,
        "phase_ms_per_step": {k: v[0] / a.steps for k, v in phases.items()}, "roofline": roofline,
        "step_roofline": step_roofline, "gpu_launches": int(launches), "clocks": clk,
        "thermo_last": {"step": st[-1], "T": T[-1], "U": U[-1], "P": P[-1]} if st else None,
        "counts": {"nlocal_rank0": nlocal, "nghost_rank0": sim.geti("nghost"), "maxneighs": sim.geti("maxneighs")},
    }


def measure_e2e(a, D, sim, ctx):
    """The same neighbor cycle through the C ABI with HOST buffers (pinned): upload, ghosts + lists + forces as the
    reference's main() does after setup (ref/ljs.cpp:445-459), 20 MD steps, download."""
    import ctypes as C

    import numpy as np  # noqa: F401

    from minimd_b200._lib import ThermoSample, check
    torch = D.torch
    s = 8 if a.precision == "f64" else 4
    natoms = sim.geti("natoms")
    nlocal = sim.geti("nlocal")
    cap = int(nlocal * 1.1) + 4096
    hx = torch.empty((cap, 3), dtype=torch.float64 if s == 8 else torch.float32).pin_memory().numpy()
    hv = torch.empty((cap, 3), dtype=torch.float64 if s == 8 else torch.float32).pin_memory().numpy()
    ht = torch.empty((cap,), dtype=torch.int32).pin_memory().numpy()
    lib = ctx.lib
    vp = lambda arr: arr.ctypes.data_as(C.c_void_p)  # noqa: E731
    half, gn = a.half_neigh, (a.ghost_newton if a.half_neigh else 0)
    params = sim.run_params(MD_STEPS_PER_STEP)
    samples = (ThermoSample * 4)()
    nsamp = C.c_int()

    def download(n):
        check(lib.mmd_atom_download(ctx.h, vp(hx), None, None, vp(ht), 0, n, 3))
        check(lib.mmd_atom_download(ctx.h, None, vp(hv), None, None, 0, n, 3))

    def cycle(n):
        check(lib.mmd_atom_upload(ctx.h, vp(hx), vp(hv), vp(ht), n, 3))
        check(lib.mmd_comm_exchange(ctx.h))
        check(lib.mmd_comm_borders(ctx.h))
        mx, tot = C.c_int(0), C.c_longlong()
        check(lib.mmd_neigh_build(ctx.h, half, gn, C.byref(mx), C.byref(tot)))
        if a.force == "lj":
            check(lib.mmd_force_lj_compute(ctx.h, half, gn, 0, None, None))
        else:
            check(lib.mmd_force_eam_compute(ctx.h, half, 0, None, None))
        if half and gn:
            check(lib.mmd_comm_reverse_communicate(ctx.h))
        check(lib.mmd_run(ctx.h, C.byref(params), samples, 4, C.byref(nsamp), None))
        n2 = ctx.counts()[0]
        download(n2)
        return n2

    n_now = ctx.counts()[0]
    download(n_now)
    n_now = cycle(n_now)          # warm-up (allocations, staging buffers)
    D.barrier()
    h2d = d2h = 0
    t0 = time.perf_counter()
    reps = max(1, min(a.steps, 10))
    for _ in range(reps):
        h2d += n_now * (2 * 3 * s + 4)
        n_now = cycle(n_now)
        d2h += n_now * (2 * 3 * s + 4) + 4 * 32
    D.barrier()
    dt = D.max(time.perf_counter() - t0)
    return {"value": natoms * MD_STEPS_PER_STEP * reps / dt / 1e6, "unit": UNIT,
            "h2d_bytes_per_step": int(D.sum(h2d) / reps), "d2h_bytes_per_step": int(D.sum(d2h) / reps),
            "steps": reps, "ms_per_step": 1e3 * dt / reps,
            "path": "mmd_atom_upload(host x,v,type) -> exchange -> borders -> neigh_build -> force -> reverse -> "
                    "mmd_run(20 MD steps) -> mmd_atom_download(host x,v,type)"}


OTHER_CONFIGS = {
    # BASELINE.json configs[2] and configs[3]: parity is covered by tests/; these lines make their speed driver-visible
    "lj_full_f32": dict(force="lj", size=80, half_neigh=0, ghost_newton=0, precision="f32", styles=(("full", 0, 0),)),
    "eam_full_f64": dict(force="eam", size=64, half_neigh=0, ghost_newton=0, precision="f64", styles=(("full", 0, 0),)),
}


def own_arm(a, n_gpus, rank, local_rank):
    import copy

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    D = Dist(n_gpus, rank, local_rank)

    # ---- the other BASELINE.json single-GPU configurations first (short runs; the headline stays last and unchanged) ----
    others = {}
    if n_gpus == 1 and not a.no_other and a.force == "lj" and a.half_neigh == 1 and a.precision == "f64":
        for name, oc in OTHER_CONFIGS.items():
            b = copy.copy(a)
            b.force, b.size, b.half_neigh, b.ghost_newton, b.precision = oc["force"], oc["size"], oc["half_neigh"], oc["ghost_newton"], oc["precision"]
            b.steps, b.warmup = max(3, min(a.steps, 6)), 3
            try:
                sim, ctx = make_sim(b, D, local_rank)
                r = measure_resident(b, D, sim, ctx, local_rank, with_clocks=False)
                r["e2e"] = measure_e2e(b, D, sim, ctx)
                sim.close()
                r.update({"metric": METRIC.replace("LJ -s 80", name), "unit": UNIT, "steps": b.steps, "warmup": b.warmup,
                          "dtype": b.precision, "config": workload_config(b, 1)})
                if not a.no_cpu_baseline:
                    r["cpu_baseline"] = cpu_baseline(b, 1, MD_STEPS_PER_STEP, styles=oc["styles"])
                others[name] = r
            except Exception as e:  # noqa: BLE001
                others[name] = {"error": repr(e)}

    sim, ctx = make_sim(a, D, local_rank)
    r = measure_resident(a, D, sim, ctx, local_rank, with_clocks=(rank == 0))
    e2e = None if a.no_e2e else measure_e2e(a, D, sim, ctx)
    result = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.precision, "data": "synthetic", "config": workload_config(a, n_gpus),
        "md_steps_timed": r["md_steps_timed"], "device_ms_inside_mmd_run": r["device_ms_inside_mmd_run"],
        "phase_ms_per_step": r["phase_ms_per_step"], "roofline": r["roofline"], "step_roofline": r["step_roofline"],
        "kernel_profile": r["kernel_profile"], "e2e": e2e, "gpu_launches": r["gpu_launches"], "clocks": r["clocks"], "thermo_last": r["thermo_last"],
        "counts": r["counts"],
        "halo_transport": ("none (single rank: device-local self swaps)" if n_gpus == 1 else
                           ("peer-memory windows over NVLink (CUDA IPC), fused pack+remote store / wait+unpack kernels"
                            if ctx.query("p2p_active") else "NCCL send/recv + pack/unpack kernels")),
    }
    if others:
        result["other_configs"] = others
    if rank == 0 and n_gpus == 1 and not a.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline(a, 1, MD_STEPS_PER_STEP)
    elif rank == 0:
        result["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(result), flush=True)
    sim.close()
    D.close()


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
    ap.add_argument("--no-other", dest="no_other", action="store_true", help="skip the other_configs legs (FP32 full list, EAM)")
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
