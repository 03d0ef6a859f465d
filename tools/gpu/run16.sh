mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile_force or isolated or fused_paths or time_loop or per_type or unusual" 2>&1 | tail -5 > gpurun_out/r2_t16.log
cat gpurun_out/r2_t16.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_bench16.json 2> gpurun_out/r2_bench16.err
cp minimd_b200/lib/libminimd_b200.so /tmp/keep.so
cp tools/gpu/libminimd_b200_prof.so minimd_b200/lib/libminimd_b200.so
for d in 1 0; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other --opt kernel_profile=1 --opt tile_dealt=$d > gpurun_out/r2_stage16_dealt$d.json 2> gpurun_out/r2_stage16_dealt$d.err
done
cp /tmp/keep.so minimd_b200/lib/libminimd_b200.so
python - <<'PY'
import json
for f in ("r2_bench16","r2_stage16_dealt1","r2_stage16_dealt0"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d.get("kernel_profile"))
    except Exception as e:
        print("ERR", e, open(f"gpurun_out/{f}.err").read()[-2000:])
PY
