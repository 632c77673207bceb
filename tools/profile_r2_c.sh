#!/bin/bash
# Round-2 profiling recipe, part c (one GPU): re-capture of the update kernels after the session-3 changes (hidden-16 and
# hidden-64 tcgen05 kernels) -- launch list + `ncu --set full` of the headline command, `ncu --set full` of the hidden-64 line.
set -u
O=gpurun_out
CMD="python bench.py --quick --steps 2 --warmup 3"
$CMD > $O/r2_plain_headline.log 2>&1 || { tail -5 $O/r2_plain_headline.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_headline.csv $CMD > $O/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ppo_table_umma16|epoch_table" -s 9 -c 3 -o $O/r2_prof_headline $CMD > $O/r2_ncu_full_headline.log 2>&1
CMD64="python bench.py --quick --domain armman --arms 1480 --envs 1024 --horizon 10 --hidden 64 --steps 2 --warmup 3"
$CMD64 > $O/r2_plain_h64.log 2>&1 || { tail -5 $O/r2_plain_h64.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"ppo_table_umma64" -s 6 -c 2 -o $O/r2_prof_h64 $CMD64 > $O/r2_ncu_full_h64.log 2>&1
ls -la $O | grep -E "r2_prof|r2_launches"
