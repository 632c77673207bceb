#!/bin/bash
# headline stage times for the tree library and every tools/bin/librmab_<name>.so named on the command line
for rep in 1 2; do
for v in tree "$@"; do
  if [ $v = tree ]; then unset RMAB_LIB_PATH; else export RMAB_LIB_PATH=$PWD/tools/bin/librmab_$v.so; fi
  python bench.py --quick --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],4), round(d['stage_ms_per_step']['ppo_update'],4))"
done; done
unset RMAB_LIB_PATH
