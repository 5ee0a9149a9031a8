#!/bin/bash
# tools/build_variant.sh <name> <nvcc -D flags...>: libpe_b200.so with pe_kernels_fused3.cu compiled under extra flags ->
# tmp_variants/libpe_b200_<name>.so (git-ignored; travels to the GPU box, where an A/B script copies it over the in-tree library)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p tmp_variants lives_b200/build
python -m lives_b200.build > /dev/null
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -c -o tmp_variants/f3_$NAME.o lives_b200/csrc/pe_kernels_fused3.cu 2>&1 | grep -i "error" || true
OBJS=$(ls lives_b200/build/*.o | grep -v pe_kernels_fused3.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -o tmp_variants/libpe_b200_$NAME.so $OBJS tmp_variants/f3_$NAME.o
rm tmp_variants/f3_$NAME.o
echo built tmp_variants/libpe_b200_$NAME.so
