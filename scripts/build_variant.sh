#!/bin/bash
# Experiment helper: libegn variant with extra -D flags on egn_fused.cu only (other objects reused from the main build).
# Usage: scripts/build_variant.sh NAME "-DFU_X=1 ..."   ->  egonerf_b200/variants/libegn_NAME.so  (select with EGN_B200_LIB)
set -e
cd "$(dirname "$0")/../egonerf_b200/csrc"
make -j8 >/dev/null
mkdir -p ../variants /tmp/egn_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -O2 $2 -c egn_fused.cu -o /tmp/egn_variants/egn_fused_$1.o
OBJS=$(ls *.o | grep -v egn_fused.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libegn_$1.so $OBJS /tmp/egn_variants/egn_fused_$1.o -lcudart
cuobjdump -res-usage /tmp/egn_variants/egn_fused_$1.o | grep -A1 "fused_fine_kernelILb1" | grep -o "REG:[0-9]* STACK:[0-9]*"
