mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_t20.log
cat gpurun_out/r2_t20.log
for o in 1 0; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other --opt fuse_ghosts=$o > gpurun_out/r2_bench20_g$o.json 2> gpurun_out/r2_bench20_g$o.err
for sz in 8 32; do
timeout 300 python bench.py --size $sz --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-other --opt fuse_ghosts=$o > gpurun_out/r2_bench20_s${sz}_g$o.json 2> gpurun_out/r2_bench20_s${sz}_g$o.err
done
done
python - <<'PY'
import json
for f in ("r2_bench20_g1","r2_bench20_g0","r2_bench20_s8_g1","r2_bench20_s8_g0","r2_bench20_s32_g1","r2_bench20_s32_g0"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],4), d["phase_ms_per_step"], d["gpu_launches"])
    except Exception as e:
        print("ERR", f, e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
