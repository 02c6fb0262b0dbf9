#!/bin/bash
# build a variant of libngp.so next to the library (kernel experiments; selected with NGP_LIBRARY=): tools/build_variant.sh <name> <nvcc flags...>
set -e
name=$1; shift
d=neuro_genetic_pong_self_play_b200
mkdir -p variants/$name
for f in $d/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $f -o variants/$name/$(basename $f).o &
done
wait
nvcc -shared -o variants/libngp_$name.so variants/$name/*.o -gencode arch=compute_100a,code=sm_100a
echo variants/libngp_$name.so
