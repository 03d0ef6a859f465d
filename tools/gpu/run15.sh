mkdir -p gpurun_out
for d in 1 0; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other --opt kernel_profile=1 --opt tile_dealt=$d > gpurun_out/r2_stage_dealt$d.json 2> gpurun_out/r2_stage_dealt$d.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err
python - <<'PY'
import json
for f in ("r2_stage_dealt1","r2_stage_dealt0","r2_bench15"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["phase_ms_per_step"], d["roofline"]["avg_launch_ms"], d.get("kernel_profile"))
    except Exception as e:
        print("ERR", e, open(f"gpurun_out/{f}.err").read()[-2000:])
PY
