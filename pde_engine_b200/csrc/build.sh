#!/bin/bash
# Build libpde_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177"
OUT=../libpde_b200.so
# --jump-table-density: the interpreter's dense micro-op switch becomes one indexed branch (brx.idx)
$NVCC $FLAGS --jump-table-density=${PDE_JTD:-25} ${PDE_PTXAS_V:+-Xptxas -v} -c pde_b200.cu -o pde_b200.o &
$NVCC $FLAGS --jump-table-density=${PDE_JTD:-25} ${PDE_PTXAS_V:+-Xptxas -v} -c program.cu -o program.o &
$NVCC $FLAGS ${PDE_PTXAS_V:+-Xptxas -v} -c enumerate.cu -o enumerate.o &
$NVCC $FLAGS -x cu -c compiler.cpp -o compiler.o &
$NVCC $FLAGS -c order.cu -o order.o &
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT pde_b200.o program.o enumerate.o compiler.o order.o -lcudart
echo "built $(realpath $OUT)"
