#!/bin/bash
# A/B builds of libf4la_b200.so with different compile-time knobs of the kernels.
# usage: tools/build_variants.sh name "-DF4LA_FLOW_OCC=3" [name2 "flags2" ...]
# result: msolve_b200/csrc/variants/libf4la_b200_<name>.so  (select with $F4LA_B200_LIB)
set -e
cd "$(dirname "$0")/../msolve_b200/csrc"
mkdir -p variants
build_one() {
  name=$1; flags=$2
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a \
      -Xcompiler -fPIC,-Wall,-Wno-unused-function,-fopenmp -I../../include $flags -Xptxas -v \
      -c f4la_device.cu -o variants/dev_$name.o 2> variants/ptxas_$name.log
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libf4la_b200_$name.so \
      variants/dev_$name.o f4la_neogb.o -lcudart_static -ldl -lpthread -lrt -l:libgomp.so.1
  rm -f variants/dev_$name.o
  echo "$(grep -A2 'k_pull_flowILi2ELb1' variants/ptxas_$name.log | grep -E 'spill|Used' | tr '\n' ' ') <- $name"
}
while [ $# -ge 2 ]; do build_one "$1" "$2" & shift 2; done
wait
