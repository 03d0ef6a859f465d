mkdir -p gpurun_out
cp minimd_b200/lib/libminimd_b200.so /tmp/keep.so
cp tools/gpu/libminimd_b200_prof.so minimd_b200/lib/libminimd_b200.so
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other --opt kernel_profile=1 > gpurun_out/r2_stage17_ldgsts_prof.json 2> gpurun_out/r2_stage17.err
cp tools/gpu/libminimd_b200_ldgsts.so minimd_b200/lib/libminimd_b200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile_force or time_loop" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_bench17_ldgsts.json 2> gpurun_out/r2_bench17.err
cp /tmp/keep.so minimd_b200/lib/libminimd_b200.so
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_bench17_bulk.json 2> gpurun_out/r2_bench17b.err
python - <<'PY'
import json
for f in ("r2_stage17_ldgsts_prof","r2_bench17_ldgsts","r2_bench17_bulk"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d.get("kernel_profile"))
    except Exception as e:
        print("ERR", f, e)
PY
