#!/bin/bash
# Experiment variant of the prediction kernels for 2..9 atoms (extra -D flags reach kreg_predict_inst.cu and kreg_predict.cu):
#   bash tools/build_variant_predict.sh NAME -DSOME_KNOB=4   ->  tools/bin/lib_NAME.so
set -e
name=$1; shift
cd "$(dirname "$0")/../mlatom_b200/csrc"
mkdir -p build/var ../../tools/bin
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ \
  -DKREG_INST=0 "$@" -c kreg_predict_inst.cu -o build/var/inst0_$name.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ \
  "$@" -c kreg_predict.cu -o build/var/predict_$name.o
objs=$(ls build/*.o | grep -v -e kreg_predict_inst0.o -e kreg_predict.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/bin/lib_$name.so $objs build/var/inst0_$name.o build/var/predict_$name.o -lcusolver -lcublas -ccbin /usr/bin/g++
echo built tools/bin/lib_$name.so
