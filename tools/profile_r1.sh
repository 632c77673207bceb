#!/bin/bash
# Round-1 profiling recipe (run under gpurun, one GPU): plain run, launch list, full capture of the dominant kernels.
set -u
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-whittle --no-armman"
$CMD > gpurun_out/plain_final.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launches_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ppo_table_umma16|epoch_table" -s 9 -c 3 -o gpurun_out/prof_final $CMD > gpurun_out/ncu_full_final.log 2>&1
ls -la gpurun_out | tail -6
