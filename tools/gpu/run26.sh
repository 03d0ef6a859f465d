mkdir -p gpurun_out
python tools/gpu/det_check.py 60 2>&1 | tail -4
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench26.json 2> gpurun_out/r2_bench26.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench26.json") if l.startswith("{")][-1])
print(round(d["value"],1), round(d["ms_per_step"],4), d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"])
for k,v in d.get("other_configs",{}).items(): print(k, round(v["value"],1), v["ms_per_step"], v["phase_ms_per_step"], v["roofline"]["frac"])
PY
