mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_t30.log
cat gpurun_out/r2_t30.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2_bench30.json 2> gpurun_out/r2_bench30.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench30_ref.json 2> gpurun_out/r2_bench30_ref.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench30.json") if l.startswith("{")][-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"].get("fp64_frac"), d["roofline"]["traffic"], round(d["e2e"]["value"],1), round(d["cpu_baseline"]["value"],2), d["clocks"])
    for k,v in d.get("other_configs",{}).items():
        print(k, v.get("error") or (round(v["value"],1), v["phase_ms_per_step"], v["roofline"]["frac"], v["roofline"]["traffic"], round(v["e2e"]["value"],1), round(v.get("cpu_baseline",{}).get("value",0),2)))
    r=json.loads([l for l in open("gpurun_out/r2_bench30_ref.json") if l.startswith("{")][-1])
    print("ref", r["value"], r["steps"], r["warmup"], r["ms_per_step"])
except Exception as e:
    print("ERR", e, open("gpurun_out/r2_bench30.err").read()[-3000:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/r2_launches30.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_l30.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/r2_launches30_eam.csv python bench.py --force eam --size 64 --half_neigh 0 --ghost_newton 0 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_l30e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eam_dealt -s 20 -c 4 -o gpurun_out/r2_prof_eam30 python bench.py --force eam --size 64 --half_neigh 0 --ghost_newton 0 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu30e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_rows_deal" -s 1 -c 1 -o gpurun_out/r2_prof_deal30 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu30d.log 2>&1
ls gpurun_out/*30*
