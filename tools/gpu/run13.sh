mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_t13.log
cat gpurun_out/r2_t13.log
for o in 1 0; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other --opt tile_pair_build=$o > gpurun_out/r2_bench13_$o.json 2> gpurun_out/r2_bench13_$o.err
done
python - <<'PY'
import json
for o in (1,0):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2_bench13_{o}.json") if l.startswith("{")][-1])
        print(o, d["value"], d["ms_per_step"], d["phase_ms_per_step"])
    except Exception as e:
        print("ERR", e, open(f"gpurun_out/r2_bench13_{o}.err").read()[-3000:])
PY
