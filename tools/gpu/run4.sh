mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_t4.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other --opt tile_lane_build=0 > gpurun_out/r2_bench4_warpbuild.json 2> gpurun_out/r2_bench4_warpbuild.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"neigh_build_lane|tile_rows_deal|xs_fill" -s 2 -c 4 -o gpurun_out/r2_prof_build4 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu4.log 2>&1
./tools/microbench/fp64_fma_bench > gpurun_out/r2_fp64_peak.json 2>&1
cat gpurun_out/r2_t4.log
cat gpurun_out/r2_fp64_peak.json
python - <<'PY'
import json
for k in ("","_warpbuild"):
    try:
        d=json.load(open(f"gpurun_out/r2_bench4{k}.json"))
        print(k, d["value"], d["ms_per_step"], d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"].get("fp64_frac"))
    except Exception as e:
        print(k, "ERR", e, open(f"gpurun_out/r2_bench4{k}.err").read()[-2000:])
PY
