mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile_force or isolated or fused_paths or time_loop or per_type" 2>&1 | tail -15 > gpurun_out/r2_t3.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench3_dealt1.json 2> gpurun_out/r2_bench3_dealt1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"force_lj_dealt|tile_rows_deal" -s 25 -c 2 -o gpurun_out/r2_prof_dealt3 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu3.log 2>&1
cat gpurun_out/r2_t3.log
python - <<'PY'
import json
for k in ("1",):
    try:
        d=json.load(open(f"gpurun_out/r2_bench3_dealt{k}.json"))
        print(k, d["value"], d["ms_per_step"], d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["clocks"])
    except Exception as e:
        print(k, "ERR", e, open(f"gpurun_out/r2_bench3_dealt{k}.err").read()[-2000:])
PY
