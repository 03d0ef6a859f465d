mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -x -q -m gpu -k "eam or EAM" 2>&1 | tail -3
timeout 600 python bench.py --force eam --size 64 --half_neigh 0 --ghost_newton 0 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_bench28.json 2> gpurun_out/r2_bench28.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench28.json") if l.startswith("{")][-1])
print(round(d["value"],1), round(d["ms_per_step"],4), d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"])
PY
