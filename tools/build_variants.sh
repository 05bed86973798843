#!/bin/bash
# A/B builds of the TMA kernel TU with different sampler switches: bayesian_torch_b200/build/variants/libbtb200_<name>.so
# (select at run time with BT_LIB_VARIANT=<path>); needs an up-to-date normal build for the other objects
set -e
cd "$(dirname "$0")/.."
B=bayesian_torch_b200/build; V=$B/variants; mkdir -p $V
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
build_one() {  # name, defines
  nvcc $FLAGS $2 -c bayesian_torch_b200/csrc/bt_tma.cu -o $V/bt_tma_$1.o
  objs=$(ls $B/*.o | grep -v bt_tma.o)
  nvcc -shared -o $V/libbtb200_$1.so $objs $V/bt_tma_$1.o -gencode arch=compute_100a,code=sm_100a -cudart static
  rm -f $V/bt_tma_$1.o
}
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  build_one "$name" "$defs" &
done
wait
ls -la $V
