#!/bin/bash
# Experiment variant of the prediction kernels (extra -D flags reach kreg_predict_inst.cu and kreg_predict.cu):
#   INST=3 bash tools/build_variant_predict.sh NAME -DSOME_KNOB=4   ->  tools/bin/lib_NAME.so
# INST selects the size range that is rebuilt (0: 2-9 atoms, 1: 10-15, 2: 16-20, 3: 21-22, 4: 23-24; default 0).
set -e
name=$1; shift
inst=${INST:-0}
cd "$(dirname "$0")/../mlatom_b200/csrc"
mkdir -p build/var ../../tools/bin
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ \
  -DKREG_INST=$inst "$@" -c kreg_predict_inst.cu -o build/var/inst${inst}_$name.o
objs=$(ls build/*.o | grep -v -e kreg_predict_inst$inst.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/bin/lib_$name.so $objs build/var/inst${inst}_$name.o -lcusolver -lcublas -ccbin /usr/bin/g++
echo built tools/bin/lib_$name.so
