#!/bin/bash
# Round-1 profiling recipe for the hidden-64 (ARMMAN shard) bench line: plain run, launch list, full capture of the
# tcgen05 update kernels.  Run under gpurun on one GPU.
set -u
CMD="python bench.py --domain armman --arms 12500 --envs 1024 --horizon 10 --hidden 64 --steps 2 --warmup 3 --no-cpu-baseline --no-whittle"
$CMD > gpurun_out/bench_armman_h64.json 2> gpurun_out/plain_h64.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_h64.csv $CMD > gpurun_out/ncu_launches_h64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"umma64" -s 6 -c 2 -o gpurun_out/prof_h64 $CMD > gpurun_out/ncu_full_h64.log 2>&1
ls -la gpurun_out | tail -6
