#!/bin/bash
# usage: build_variant.sh NAME -DFLAG=.. ...  -> video_transformer_b200/libvtseg_NAME.so (measurement builds of the pair scaler)
set -e
name=$1; shift
cd "$(dirname "$0")/../video_transformer_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --use_fast_math -Xcompiler -fPIC,-O2 "$@" -c -o csrc/_obj/vt_scale_pair_$name.o csrc/vt_scale_pair.cu
objs=$(ls csrc/_obj/*.o | grep -v "vt_scale_pair" | tr '\n' ' ')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -o libvtseg_$name.so $objs csrc/_obj/vt_scale_pair_$name.o -ldl
cuobjdump -res-usage csrc/_obj/vt_scale_pair_$name.o | grep -A1 "scale_pair_kernelILi3ELi6" | grep REG | awk '{print $1}'
