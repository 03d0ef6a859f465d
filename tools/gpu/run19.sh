mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python tools/multi_rank_record.py gpurun_out/r2_multi_rank_check19_n$N.json $N 2>&1 | cut -c1-260 | tail -9
timeout 600 python bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_bench19_n${N}.json 2> gpurun_out/r2_bench19_n${N}.err
python - <<PY
import json
try:
    d=json.loads([l for l in open(f"gpurun_out/r2_bench19_n${N}.json") if l.startswith("{")][-1])
    print(d["n_gpus"], d["value"], d["ms_per_step"], d["phase_ms_per_step"])
except Exception as e:
    print("ERR", e, open(f"gpurun_out/r2_bench19_n${N}.err").read()[-3000:])
PY
