mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_t21.log
cat gpurun_out/r2_t21.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench21.json 2> gpurun_out/r2_bench21.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench21.json") if l.startswith("{")][-1])
print(round(d["value"],1), round(d["ms_per_step"],4), d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"])
for k,v in d.get("other_configs",{}).items():
    print(k, v.get("error") or (round(v["value"],1), v["phase_ms_per_step"], v["roofline"]["frac"]))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_rows_deal|force_lj_dealt" -s 4 -c 3 -o gpurun_out/r2_prof_deal21 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu21.log 2>&1
