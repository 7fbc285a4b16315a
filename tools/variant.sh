#!/bin/bash
# Build an alternative copy of the library with extra nvcc flags (tuning experiments):
#   bash tools/variant.sh <name> [-DFLAG ...]   ->  build/variants/libdan_<name>.so   (use with DAN_B200_LIB=...)
set -e
NAME=$1; shift
mkdir -p build/variants
SRC=""
for s in api anchors encode postprocess mining routing vote handoff gather; do SRC="$SRC dan_b200/csrc/$s.cu"; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Iinclude -Idan_b200/csrc \
  -shared "$@" $SRC -o build/variants/libdan_$NAME.so
echo build/variants/libdan_$NAME.so
