mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other --opt kernel_profile=1 > gpurun_out/r2_stage31.json 2> gpurun_out/r2_stage31.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_stage31.json") if l.startswith("{")][-1])
print(d["value"], d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d.get("kernel_profile"))
PY
