#!/bin/bash
# rebuild msda_planar.cu with extra -D flags and relink (benchmark variants on the GPU box): tools/debug/variant.sh -DX=1 ...
cd snipper_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden "$@" -Xptxas=-v -c msda_planar.cu -o ../lib/obj/msda_planar.o 2>&1 | grep -E "registers|spill" | grep -A1 -E "planar_(fwd|bwd)_kernelILi32" | grep -E "Used|spill" | head -4
nvcc -shared -Xcompiler -fPIC -o ../lib/libmsda_b200.so ../lib/obj/*.o
