#!/bin/bash
# Builds measurement variants of libvtseg.so with parts of the pair kernel removed: video_transformer_b200/libvtseg_ablN.so
set -e
cd "$(dirname "$0")/../video_transformer_b200"
python -m video_transformer_b200.build 2>/dev/null || (cd .. && python -m video_transformer_b200.build)
for n in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --use_fast_math -Xcompiler -fPIC,-O2 -DVT_ABLATE=$n -c -o csrc/_obj/vt_scale_pair_abl$n.o csrc/vt_scale_pair.cu &
done
wait
for n in "$@"; do
  objs=$(ls csrc/_obj/*.o | grep -v "vt_scale_pair" | tr '\n' ' ')
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -o libvtseg_abl$n.so $objs csrc/_obj/vt_scale_pair_abl$n.o -ldl
done
ls -la libvtseg*.so
