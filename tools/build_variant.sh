#!/bin/bash
# Builds a VARIANT of librtds.so with extra nvcc flags for A/B runs (RTDS_LIB selects it at run time):
#   tools/build_variant.sh mb5 -DRTDS_PK_MINB=5      ->  raytracer-data-structures_b200/variants/librtds_mb5.so
#   RTDS_LIB=$PWD/raytracer-data-structures_b200/variants/librtds_mb5.so python tools/ab_render.py RTDS_HULL=1
# Only render.cu is recompiled with the flags (the other objects are the default build's); *.so is git-ignored and ships to
# the GPU box with gpurun. Delete the variants directory when done.
set -e
cd "$(dirname "$0")/../raytracer-data-structures_b200/csrc"
name=$1; shift
make -j"$(nproc)" >/dev/null
mkdir -p ../variants
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC"
nvcc $NVF "$@" -c render.cu -o /tmp/render_"$name".o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/librtds_"$name".so api.o sort.o lbvh.o median.o sah.o kd.o /tmp/render_"$name".o -lcudart
cuobjdump --dump-resource-usage /tmp/render_"$name".o 2>/dev/null | grep -A1 "render_packet_kernelILb0ELb1" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | paste - -
echo "built ../variants/librtds_$name.so"
