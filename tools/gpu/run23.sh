mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_t23.log
cat gpurun_out/r2_t23.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2_bench23.json 2> gpurun_out/r2_bench23.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench23_ref.json 2> gpurun_out/r2_bench23_ref.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench23.json") if l.startswith("{")][-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"].get("fp64_frac"), d["roofline"]["traffic"], round(d["e2e"]["value"],1), round(d["cpu_baseline"]["value"],2), d["clocks"])
    for k,v in d.get("other_configs",{}).items():
        print(k, v.get("error") or (round(v["value"],1), v["phase_ms_per_step"], v["roofline"]["frac"], v["roofline"]["traffic"], round(v["e2e"]["value"],1), round(v.get("cpu_baseline",{}).get("value",0),2)))
    r=json.loads([l for l in open("gpurun_out/r2_bench23_ref.json") if l.startswith("{")][-1])
    print("ref", r["value"], r["steps"], r["warmup"], r["ms_per_step"])
except Exception as e:
    print("ERR", e, open("gpurun_out/r2_bench23.err").read()[-3000:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"force_lj_dealt" -s 25 -c 2 -o gpurun_out/r2_prof_force23 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu23a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_rows_deal|neigh_build_tile3" -s 2 -c 2 -o gpurun_out/r2_prof_build23 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu23b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/r2_launches23.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_l23.log 2>&1
