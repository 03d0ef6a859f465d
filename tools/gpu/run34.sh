mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python - <<'PY'
import sys, json, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
from test_gpu_multi import CASES, launch
n_avail = torch.cuda.device_count()
out = []
for k, (n, extra) in enumerate(CASES):
    if n == n_avail and ("--golden" in extra or "eam" in extra):
        r = launch(n, extra, 29800 + k)
        out.append({"ranks": n, "args": [str(e) for e in extra], "result": r})
        print(json.dumps(out[-1])[:500])
json.dump(out, open(f"gpurun_out/r2_multi_final_n{n_avail}.json", "w"), indent=1)
PY
timeout 600 python bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_bench34_n${N}.json 2> gpurun_out/r2_bench34_n${N}.err
python - <<PY
import json
try:
    d=json.loads([l for l in open(f"gpurun_out/r2_bench34_n${N}.json") if l.startswith("{")][-1])
    print(d["n_gpus"], d["value"], d["ms_per_step"], d["phase_ms_per_step"])
except Exception as e:
    print("ERR", e, open(f"gpurun_out/r2_bench34_n${N}.err").read()[-3000:])
PY
