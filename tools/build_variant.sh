#!/bin/bash
# usage: tools/build_variant.sh <name> <file.cu> [-DFLAG=..]...  -> tools/bin/librmab_<name>.so : the tree library with one
# translation unit rebuilt under extra defines (A/B experiments; pick it up with RMAB_LIB_PATH).
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/.."
python -m robustrmab_b200.build > /dev/null
mkdir -p tools/bin /tmp/variant_$name
obj=/tmp/variant_$name/$(basename $src .cu).o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c robustrmab_b200/csrc/$src -o $obj
objs=$(ls robustrmab_b200/build/*.o | grep -v "/$(basename $src .cu).o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/bin/librmab_$name.so $objs $obj -lcudart
echo tools/bin/librmab_$name.so
