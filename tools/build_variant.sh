#!/bin/bash
# Build a variant of libbias_cuda.so with extra -D switches (kernel experiments):
#   tools/build_variant.sh NAME -DBIAS_EXP_CVT=1 ...  ->  bias.jl_b200/csrc/_exp/libbias_NAME.so
# Run with BIAS_CUDA_LIB=bias.jl_b200/csrc/_exp/libbias_NAME.so python bench.py ...
set -e
name=$1; shift
cd "$(dirname "$0")/../bias.jl_b200/csrc"
mkdir -p _exp/$name
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $*"
nvcc $FLAGS -c model.cu -o _exp/$name/model.o &
nvcc $FLAGS -c hdp.cu -o _exp/$name/hdp.o &
nvcc $FLAGS -c comm.cu -o _exp/$name/comm.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _exp/libbias_$name.so _exp/$name/model.o _exp/$name/hdp.o _exp/$name/comm.o -Xlinker --exclude-libs,ALL -lcudart_static -ldl -lrt -lpthread
rm -rf _exp/$name
echo built _exp/libbias_$name.so
