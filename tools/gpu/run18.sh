mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 900 python tools/multi_rank_record.py gpurun_out/r2_multi_rank_check_n$N.json $N 2>&1 | cut -c1-300 | tail -5
for s in 1 0; do
timeout 600 python bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --opt split_force=$s > gpurun_out/r2_bench18_n${N}_split$s.json 2> gpurun_out/r2_bench18_n${N}_split$s.err
done
python - <<PY
import json
for s in (1,0):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2_bench18_n${N}_split{s}.json") if l.startswith("{")][-1])
        print(s, d["n_gpus"], d["value"], d["ms_per_step"], d["phase_ms_per_step"], d["halo_transport"][:30])
    except Exception as e:
        print("ERR", e, open(f"gpurun_out/r2_bench18_n${N}_split{s}.err").read()[-3000:])
PY
