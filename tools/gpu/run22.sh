mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "golden" 2>&1 | tail -4
python - <<'PY'
import sys, json, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
from test_gpu_multi import CASES, launch
n_avail = torch.cuda.device_count()
out = []
for k, (n, extra) in enumerate(CASES):
    if "--golden" in extra and n == n_avail:
        r = launch(n, extra, 29800 + k)
        out.append({"ranks": n, "args": [str(e) for e in extra], "result": r})
        print(json.dumps(out[-1])[:600])
json.dump(out, open(f"gpurun_out/r2_multi_golden_n{n_avail}.json", "w"), indent=1)
PY
