#!/bin/bash
# Builds a variant of libhbt_b200.so with extra -D switches (A/B runs and control experiments):
#   scripts/build_variant.sh redoff -DHBT_DBG_RED=1
# -> hadronic_afterburner_toolkit_b200/variants/libhbt_b200_redoff.so, loaded with HBT_B200_LIB=<path>.
set -e
name=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
pkg=$here/hadronic_afterburner_toolkit_b200
out=$pkg/variants
mkdir -p $out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -DHBT_HAVE_V2 --extended-lambda \
    -Xcompiler "-O2 -fPIC -ffp-contract=off -fno-fast-math" "$@" -Xptxas -v -c $pkg/csrc/hbt_b200.cu -o $out/hbt_b200_$name.o 2> $out/ptxas_$name.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libhbt_b200_$name.so $out/hbt_b200_$name.o $pkg/csrc/hbt_bf.o $pkg/csrc/hbt_host.o $pkg/csrc/hbt_reader.o $pkg/csrc/hbt_inflate.o -ldl -lz -lpthread
rm -f $out/hbt_b200_$name.o
grep -A2 "hbt_pairs_v3" $out/ptxas_$name.log | grep -E "Used" | sort | uniq -c
