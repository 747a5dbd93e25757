#!/bin/bash
# Build an A/B variant of the library: tools/build_variant.sh <name> <extra nvcc flags for ba.cu...>  ->  build/libcmos_<name>.so
set -e
NAME=$1; shift
cd "$(dirname "$0")/../ceres_mono_orb_slam2_b200/csrc"
O=../../build/obj
mkdir -p $O
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-ffp-contract=off "$@" -c ba.cu -o $O/ba_$NAME.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/libcmos_$NAME.so $O/common.o $O/orb.o $O/match.o $O/match_kf.o $O/mappoint.o $O/voc.o $O/track.o $O/ba_$NAME.o -lcudart_static -lpthread -ldl -lrt
echo build/libcmos_$NAME.so
