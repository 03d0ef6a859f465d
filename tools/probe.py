#!/usr/bin/env python
"""GPU probe: per-kernel timings (CUDA events via torch on the context's stream), TPA sweep and a
timed mmd_run at a given size.  Development aid; bench.py is the judged measurement."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from helpers import context_from_oracle, run_params, thermo_from_samples
from oracle.oracle import Config, Oracle


def timeit(ctx, fn, reps=5, warm=1):
    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(warm):
        fn()
    ctx.sync()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def profile_mode(a):
    """One launch of each interesting kernel variant (run under `ncu -k regex:...`)."""
    from helpers import geometry_of
    gn = 1 if a.force == "lj" else 0
    cfg = Config(nx=a.s, ny=a.s, nz=a.s, force=a.force, halfneigh=1, ghost_newton=gn)
    o = Oracle(cfg, a.prec)
    o.run(20)                      # a melted, re-neighbored state (sorted atoms), not the perfect lattice
    c = context_from_oracle(o)
    c.exchange(); c.borders()
    key = "lj_threads_per_atom" if a.force == "lj" else "eam_threads_per_atom"
    for half, g, tpas in ((1, gn, (2, 4, 8)), (0, 0, (2, 4, 8))):
        o.seti("halfneigh", half); o.seti("ghost_newton", g); o.call("neighbor_setup")
        c.neigh_setup(geometry_of(o), o.stencil(), o.rvec("cutneighsq", cfg.ntypes ** 2))
        c.build(half, g, 100)
        for tpa in tpas:
            c.set_option(key, tpa)
            for ev in (0, 1):
                if a.force == "lj":
                    c.lib.mmd_force_lj_compute(c.h, half, g, ev, None, None)
                else:
                    c.lib.mmd_force_eam_compute(c.h, half, ev, None, None)
        c.sync()
    c.initial_integrate(0.0, 0.0); c.final_integrate(0.0); c.communicate(); c.reverse_communicate()
    c.sort(); c.borders(); c.sync()
    print("profile mode done", c.launches)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-s", type=int, default=80)
    ap.add_argument("--force", default="lj")
    ap.add_argument("--prec", default="f64")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--out", default="gpurun_out/probe.json")
    ap.add_argument("--profile", action="store_true", help="few launches only: meant to run under ncu")
    a = ap.parse_args()
    if a.profile:
        return profile_mode(a)
    res = {"size": a.s, "force": a.force, "prec": a.prec}
    gn = 1 if a.force == "lj" else 0
    cfg = Config(nx=a.s, ny=a.s, nz=a.s, force=a.force, halfneigh=1, ghost_newton=gn)
    t0 = time.time()
    o = Oracle(cfg, a.prec)
    res["oracle_init_s"] = time.time() - t0
    natoms = o.geti("natoms")
    c = context_from_oracle(o)
    c.exchange()
    c.borders()
    res["nghost"] = c.nghost
    for half, g in ((1, gn), (0, 0)):
        key = "half" if half else "full"
        # stencil differs between half+gn and full: re-setup the oracle geometry for the list style
        o.seti("halfneigh", half); o.seti("ghost_newton", g); o.call("neighbor_setup")
        from helpers import geometry_of
        c.neigh_setup(geometry_of(o), o.stencil(), o.rvec("cutneighsq", cfg.ntypes ** 2))
        res[f"build_{key}_ms"] = timeit(c, lambda: c.build(half, g, 100), reps=3)
        res[f"total_neigh_{key}"] = c.query("total_neigh")
        for tpa in (1, 2, 4, 8, 16, 32):
            c.set_option("lj_threads_per_atom" if a.force == "lj" else "eam_threads_per_atom", tpa)
            for ev in (0, 1):
                if a.force == "lj":
                    fn = lambda: c.lib.mmd_force_lj_compute(c.h, half, g, ev, None, None)
                else:
                    fn = lambda: c.lib.mmd_force_eam_compute(c.h, half, ev, None, None)
                res[f"force_{key}_tpa{tpa}_ev{ev}_ms"] = timeit(c, fn)
        print(json.dumps({k: v for k, v in res.items() if key in k}), flush=True)
    dt, dtf = o.getr("dt"), o.getr("dtforce")
    res["initial_ms"] = timeit(c, lambda: c.initial_integrate(0.0, 0.0))
    res["final_ms"] = timeit(c, lambda: c.final_integrate(0.0))
    res["communicate_ms"] = timeit(c, lambda: c.communicate())
    res["reverse_ms"] = timeit(c, lambda: c.reverse_communicate())
    res["borders_ms"] = timeit(c, lambda: c.borders(), reps=3)
    res["sort_ms"] = timeit(c, lambda: (c.sort(), c.borders()), reps=3)
    res["binatoms_ms"] = timeit(c, lambda: c.binatoms(-1, 8), reps=3)
    # timed run from a fresh, reference-identical start (half list, default settings)
    o.seti("halfneigh", 1); o.seti("ghost_newton", gn); o.call("neighbor_setup")
    c2 = context_from_oracle(o)
    c2.exchange(); c2.borders(); c2.build(1, gn, 100)
    c2.set_option("lj_threads_per_atom", 8)
    samples, ms = c2.run(run_params(o, a.steps))
    res["run_ms"] = ms
    res["matom_steps_per_s"] = natoms * a.steps / ms / 1e3
    res["thermo"] = thermo_from_samples(o, samples)
    res["launches"] = c2.launches
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
