"""Run-to-run reproducibility of the time loop (same options twice, then fuse_halo 0 vs 1)."""
import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as t
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
runs = {}
for name, opts in (("a1", dict(fuse_halo=0, fuse_force=0, fuse_integrate=0)), ("a2", dict(fuse_halo=0, fuse_force=0, fuse_integrate=0)),
                   ("b1", dict(fuse_halo=1, fuse_force=0, fuse_integrate=0)), ("b2", dict(fuse_halo=1, fuse_force=0, fuse_integrate=0)),
                   ("c1", dict(fuse_halo=1, fuse_force=0, fuse_integrate=0, tile_dealt=0))):
    runs[name] = t._run_with(opts, "f64", steps)
for p, q in (("a1", "a2"), ("b1", "b2"), ("a1", "b1"), ("a1", "c1")):
    dx = np.abs(runs[p][1] - runs[q][1])
    print(p, q, "max|dx|", dx.max(), "atoms differing", int((dx.max(axis=1) > 0).sum()))
