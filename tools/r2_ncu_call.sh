#!/usr/bin/env bash
# Round-2 ncu evidence of the default path at c2: launch list + one full capture of the two blend kernels and the sort kernels.
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches_ncu.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/r2_ncu_launch_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 8 -c 2 -f -o gpurun_out/r2_render_default \
    python bench.py --steps 1 --warmup 4 --no-cpu-baseline --no-c4 > gpurun_out/r2_ncu_full_bench.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/r2_render_default.ncu-rep gpurun_out/r2_final_launches_ncu.csv
true
