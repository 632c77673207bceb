#!/bin/bash
# Round-2 profiling recipe, part b (one GPU): the hidden-64 update kernels (both launches) and the launch list of one
# MA-RMABPPO epoch at a reduced arm count (kernel shares of configs[3]'s domain).
set -u
O=gpurun_out
CMD="python bench.py --quick --domain armman --arms 1480 --envs 1024 --horizon 10 --hidden 64 --steps 2 --warmup 3"
$CMD > $O/r2_plain_h64.log 2>&1 || { tail -5 $O/r2_plain_h64.log; exit 1; }
MA="python tools/ma_scale_run.py 512 256"
MA_EPOCHS=3 RMAB_MA_GRAPHS=0 $MA > $O/r2_plain_ma.log 2>&1 || { tail -5 $O/r2_plain_ma.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"ppo_table_umma64" -s 6 -c 2 -o $O/r2_prof_h64 $CMD > $O/r2_ncu_full_h64.log 2>&1
MA_EPOCHS=2 RMAB_MA_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2_launches_ma.csv $MA > $O/r2_ncu_launches_ma.log 2>&1
ls -la $O | grep -E "r2_prof_h64|r2_launches_ma"
