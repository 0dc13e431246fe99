#!/bin/bash
# Experiment helper: libegn variant with extra -D flags on ONE source file (other objects reused from the main build).
# Usage: scripts/build_variant.sh NAME "-DFU_X=1 ..." [file.cu]  ->  egonerf_b200/variants/libegn_NAME.so  (select with EGN_B200_LIB)
set -e
SRC=${3:-egn_fused.cu}
cd "$(dirname "$0")/../egonerf_b200/csrc"
make -j8 >/dev/null
mkdir -p ../variants /tmp/egn_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -O2 $2 -c $SRC -o /tmp/egn_variants/${SRC%.cu}_$1.o
OBJS=$(ls *.o | grep -v ${SRC%.cu}.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libegn_$1.so $OBJS /tmp/egn_variants/${SRC%.cu}_$1.o -lcudart
cuobjdump -res-usage /tmp/egn_variants/${SRC%.cu}_$1.o | grep -A1 "${4:-fused_fine_kernelILb1}" | grep -o "REG:[0-9]* STACK:[0-9]*"
