#!/bin/bash
# Experiment builds of the library with parts of the conv kernel compiled out (WSIS_KO bits: 1 no MMA, 2 no row-cache
# reads, 4 no TMEM operand store, 8 no gather loads, 16 no epilogue read-out, 32 no gather conversion): timing only,
# results are wrong.  Usage: tools/build_knockouts.sh 1 2 4 ...  ->  3d-wsis_b200/csrc/build/ko/libwsis_ko<bits>.so
set -e
cd "$(dirname "$0")/../3d-wsis_b200/csrc"
mkdir -p build/ko build
for ko in "$@"; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DWSIS_KO=$ko -c conv_umma.cu -o build/conv_umma_ko$ko.o
  objs=$(ls build/*.o | grep -v conv_umma)
  /usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/ko/libwsis_ko$ko.so $objs build/conv_umma_ko$ko.o -lcudart_static -lpthread -ldl -lrt
  rm build/conv_umma_ko$ko.o
done
