mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python tools/multi_rank_record.py gpurun_out/r2_multi_rank_check_n2.json 2>&1 | tail -12
timeout 600 python bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_bench7_n2.json 2> gpurun_out/r2_bench7_n2.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench7_n2.json"))
    print(d["value"], d["ms_per_step"], d["phase_ms_per_step"], d["halo_transport"])
except Exception as e:
    print("ERR", e, open("gpurun_out/r2_bench7_n2.err").read()[-3000:])
PY
