#!/bin/bash
# Build a kernel-variant library for A/B experiments on the GPU box:
#   tools/build_variant.sh NAME [extra nvcc flags...]   ->  variants/libpde_NAME.so
# Select it at run time with PDE_B200_LIB=variants/libpde_NAME.so (pde_engine_b200/_lib.py).
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/pde_engine_b200/csrc
OUT=$ROOT/variants
mkdir -p $OUT/obj_$NAME
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177"
$NVCC $FLAGS --jump-table-density=${PDE_JTD:-25} "$@" -Xptxas -v -c $SRC/pde_b200.cu -o $OUT/obj_$NAME/pde_b200.o 2> $OUT/obj_$NAME/ptxas.log &
[ -f $SRC/program.o ] || $NVCC $FLAGS --jump-table-density=${PDE_JTD:-25} -c $SRC/program.cu -o $SRC/program.o
[ -f $SRC/enumerate.o ] || $NVCC $FLAGS -c $SRC/enumerate.cu -o $SRC/enumerate.o
[ -f $SRC/compiler.o ] || $NVCC $FLAGS -x cu -c $SRC/compiler.cpp -o $SRC/compiler.o
wait
$NVCC -shared -o $OUT/libpde_$NAME.so $OUT/obj_$NAME/pde_b200.o $SRC/program.o $SRC/enumerate.o $SRC/compiler.o -lcudart
grep -A2 "validate_kernelILi0ELb0ELi2[04]ELi1ELi1ELb0" $OUT/obj_$NAME/ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
