mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -4
for s in 8 32; do for g in 0 2; do
timeout 300 python bench.py --size $s --steps 50 --warmup 10 --no-e2e --no-cpu-baseline --no-other --opt graph_steps=$g > gpurun_out/r2_graph_s${s}_g$g.json 2> gpurun_out/r2_graph_s${s}_g$g.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_graph_s${s}_g$g.json") if l.startswith("{")][-1])
    print("size $s graph $g", round(d["value"],1), round(d["ms_per_step"],4), d["phase_ms_per_step"], d["gpu_launches"])
except Exception as e:
    print("ERR", e, open("gpurun_out/r2_graph_s${s}_g$g.err").read()[-2000:])
PY
done; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_bench35.json 2> gpurun_out/r2_bench35.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench35.json") if l.startswith("{")][-1])
print(round(d["value"],1), round(d["ms_per_step"],4), d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"])
PY
