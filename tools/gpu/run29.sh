mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"force_lj_dealt" -s 25 -c 2 -o gpurun_out/r2_prof_force29 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu29a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_rows_deal|neigh_build_tile3" -s 2 -c 2 -o gpurun_out/r2_prof_build29 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu29b.log 2>&1
ls -la gpurun_out/*29*
