#!/bin/bash
# A/B of the hidden-16 update kernels under gpurun: the library in the tree against tools/bin/librmab_old.so (a saved build),
# same bench command, stage times side by side; then the parity tests that exercise the hidden-16 path.
O=gpurun_out
for i in 1 2; do
  RMAB_LIB_PATH=$PWD/tools/bin/librmab_old.so python bench.py --quick --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('old', d['ms_per_step'], d['stage_ms_per_step'])"
  python bench.py --quick --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new', d['ms_per_step'], d['stage_ms_per_step'])"
done
python -m pytest tests/test_gpu_epoch.py tests/test_gpu_kernels.py tests/test_gpu_dropin.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8
