mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"force_lj_dealt" -s 25 -c 1 -o gpurun_out/r2_prof_force32 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu32a.log 2>&1
ls -la gpurun_out/r2_prof_force32.ncu-rep
