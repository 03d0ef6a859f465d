mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_t12.log
cat gpurun_out/r2_t12.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench12.json") if l.startswith("{")][-1])
    print(d["value"], d["ms_per_step"], d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"].get("fp64_frac"), d["e2e"]["value"], d["cpu_baseline"]["value"])
    for k,v in d.get("other_configs",{}).items():
        print(k, v.get("error") or (v["value"], v["ms_per_step"], v["phase_ms_per_step"], v["roofline"]["frac"], v["e2e"]["value"], v.get("cpu_baseline",{}).get("value")))
except Exception as e:
    print("ERR", e, open("gpurun_out/r2_bench12.err").read()[-3000:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/r2_launches12.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_l12.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"force_lj_dealt" -s 25 -c 2 -o gpurun_out/r2_prof_force12 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu12a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_rows_deal|neigh_build_tile3|bin_xsort|xs_fill" -s 4 -c 4 -o gpurun_out/r2_prof_build12 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu12b.log 2>&1
