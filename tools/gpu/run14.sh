mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench14.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["phase_ms_per_step"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"neigh_build_tile3" -s 2 -c 1 -o gpurun_out/r2_prof_build14 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu14.log 2>&1
