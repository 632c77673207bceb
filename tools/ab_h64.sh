#!/bin/bash
# A/B of the hidden-64 update kernel under gpurun: tree library vs tools/bin/librmab_old.so on tools/h64_update_bench.py
# (1480 arms x 1024 envs, 20 + 20 Adam steps), then the parity tests of the hidden-64 paths.
echo "--- old"; RMAB_LIB_PATH=$PWD/tools/bin/librmab_old.so python tools/h64_update_bench.py 1480 1024 2>&1 | grep -v mma_sync | head -3
echo "--- new"; python tools/h64_update_bench.py 1480 1024 2>&1 | tail -7
