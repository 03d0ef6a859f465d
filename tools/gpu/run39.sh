mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r2_bench39.json 2> gpurun_out/r2_bench39.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench39_ref.json 2> gpurun_out/r2_bench39_ref.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench39.json") if l.startswith("{")][-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"].get("fp64_frac"), round(d["e2e"]["value"],1), round(d["cpu_baseline"]["value"],2), d["clocks"]["reasons"], d["gpu_launches"])
    for k,v in d.get("other_configs",{}).items():
        print(k, v.get("error") or (round(v["value"],1), v["roofline"]["frac"], round(v["e2e"]["value"],1), round(v.get("cpu_baseline",{}).get("value",0),2)))
    r=json.loads([l for l in open("gpurun_out/r2_bench39_ref.json") if l.startswith("{")][-1])
    print("ref", r["value"], r["steps"], r["warmup"], r["ms_per_step"])
except Exception as e:
    print("ERR", e, open("gpurun_out/r2_bench39.err").read()[-3000:])
PY
