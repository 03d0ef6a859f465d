mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/r2_launches6.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_l6.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_rows_deal|neigh_build_tile3" -s 2 -c 2 -o gpurun_out/r2_prof_build6 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > gpurun_out/r2_ncu6.log 2>&1
tail -3 gpurun_out/r2_l6.log
