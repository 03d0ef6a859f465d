mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_t8.log
cat gpurun_out/r2_t8.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench8.json") if l.startswith("{")][-1])
    print(d["value"], d["ms_per_step"], d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"].get("fp64_frac"), d["e2e"]["value"])
    for k,v in d.get("other_configs",{}).items():
        print(k, v.get("error") or (v["value"], v["ms_per_step"], v["phase_ms_per_step"], v["roofline"]["frac"], v["roofline"]["kernel"], v["e2e"]["value"]))
except Exception as e:
    print("ERR", e, open("gpurun_out/r2_bench8.err").read()[-3000:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"eam_dealt" -s 20 -c 4 -o gpurun_out/r2_prof_eam8 python bench.py --force eam --size 64 --half_neigh 0 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu8.log 2>&1
