mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bit_exact or tile_force or time_loop" 2>&1 | tail -5 > gpurun_out/r2_t5.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2_bench5_ref.json 2> gpurun_out/r2_bench5_ref.err
cat gpurun_out/r2_t5.log
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench5.json"))
    print(d["value"], d["ms_per_step"], d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"].get("fp64_frac"), d["e2e"]["value"], d["cpu_baseline"]["value"])
    for k,v in d.get("other_configs",{}).items():
        print(k, v.get("error") or (v["value"], v["ms_per_step"], v["phase_ms_per_step"], v["roofline"]["frac"], v["roofline"]["kernel"], v["e2e"]["value"], v.get("cpu_baseline",{}).get("value")))
except Exception as e:
    print("ERR", e, open("gpurun_out/r2_bench5.err").read()[-3000:])
try:
    r=json.load(open("gpurun_out/r2_bench5_ref.json"))
    print("ref", r["value"], r["steps"], r["warmup"], r["ms_per_step"], r["cpu_baseline"]["sample"])
except Exception as e:
    print("REF ERR", e, open("gpurun_out/r2_bench5_ref.err").read()[-3000:])
PY
