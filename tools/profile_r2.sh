#!/bin/bash
# Round-2 profiling recipe (run under gpurun, one GPU): plain runs first, then the launch list of the headline command and
# `ncu --set full` captures of the dominant kernels of the headline and of every stand-alone kernel bench.py quotes.
set -u
O=gpurun_out
CMD="python bench.py --quick --steps 2 --warmup 3"
$CMD > $O/r2_plain_headline.log 2>&1 || { tail -5 $O/r2_plain_headline.log; exit 1; }
python tools/kernel_probe.py > $O/r2_plain_probe.log 2>&1 || { tail -5 $O/r2_plain_probe.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_headline.csv $CMD > $O/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ppo_table_umma16|epoch_table" -s 9 -c 3 -o $O/r2_prof_headline $CMD > $O/r2_ncu_full_headline.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm3h|split16|select_hist|select_mark|env_step_table|gae_kernel|vi_small|whittle_small" -c 24 -o $O/r2_prof_kernels python tools/kernel_probe.py > $O/r2_ncu_full_kernels.log 2>&1
ls -la $O | grep r2_prof
