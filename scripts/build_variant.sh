#!/bin/bash
# Link a variant of libcrgpu.so whose score kernel (gotoh_score2.cu) is compiled with extra flags, for A/B runs on
# one GPU box: scripts/build_variant.sh r128 -DSCORE2_MAXNREG=128  ->  crispresso_b200/libcrgpu_r128.so
# (select it with CRGPU_LIB=crispresso_b200/libcrgpu_r128.so).  The default library must be built first.
set -e
name=$1; shift
cd "$(dirname "$0")/../crispresso_b200"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
nvcc $FLAGS "$@" -c -o build/gotoh_score2_$name.o csrc/gotoh_score2.cu
objs=$(ls build/*.o | grep -v "gotoh_score2")
nvcc $FLAGS -shared -cudart shared -o libcrgpu_$name.so $objs build/gotoh_score2_$name.o
echo libcrgpu_$name.so
