mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "eam or bit_exact or time_loop" 2>&1 | tail -5 > gpurun_out/r2_t9.log
cat gpurun_out/r2_t9.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench9.json") if l.startswith("{")][-1])
    print(d["value"], d["ms_per_step"], d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"].get("fp64_frac"))
    for k,v in d.get("other_configs",{}).items():
        print(k, v.get("error") or (v["value"], v["ms_per_step"], v["phase_ms_per_step"], v["roofline"]["frac"], v["roofline"]["kernel"], v["e2e"]["value"]))
except Exception as e:
    print("ERR", e, open("gpurun_out/r2_bench9.err").read()[-3000:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/r2_launches9_eam.csv python bench.py --force eam --size 64 --half_neigh 0 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_l9.log 2>&1
