#!/usr/bin/env python
"""Run every tests/test_gpu_multi.py case the visible GPUs allow and write the checker's JSON lines to one record
(committed as profiles/r2_multi_rank_check.json): python tools/multi_rank_record.py OUT.json [ranks]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from test_gpu_multi import CASES, launch  # noqa: E402

n_avail = torch.cuda.device_count()
only = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # optional: run the cases of exactly this many ranks
out = {"gpus_visible": n_avail, "gpu": torch.cuda.get_device_name(0) if n_avail else None, "cases": []}
for k, (n, extra) in enumerate(CASES):
    rec = {"ranks": n, "args": [str(e) for e in extra]}
    if only and n != only:
        continue
    if n > n_avail:
        rec["skipped"] = f"needs {n} GPUs"
    else:
        try:
            rec["result"] = launch(n, extra, 29700 + k)
        except AssertionError as e:  # the checker exited non-zero: keep what it printed
            rec["error"] = str(e)[-2000:]
    out["cases"].append(rec)
    print(json.dumps(rec), flush=True)
with open(sys.argv[1], "w") as fh:
    json.dump(out, fh, indent=1)
