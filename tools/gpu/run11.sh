mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1500 python tools/multi_rank_record.py gpurun_out/r2_multi_rank_check_n2b.json 2>&1 | cut -c1-400 | tail -14
for s in 1 0; do
timeout 600 python bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --opt split_force=$s > gpurun_out/r2_bench11_n2_split$s.json 2> gpurun_out/r2_bench11_n2_split$s.err
done
python - <<'PY'
import json
for s in (1,0):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2_bench11_n2_split{s}.json") if l.startswith("{")][-1])
        print(s, d["value"], d["ms_per_step"], d["phase_ms_per_step"])
    except Exception as e:
        print("ERR", e, open(f"gpurun_out/r2_bench11_n2_split{s}.err").read()[-3000:])
PY
