#!/bin/bash
# builds timing-experiment variants of libgsdr.so (GSDR_EXP bit mask) into build_exp/
set -e
mkdir -p build_exp
for e in "$@"; do
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DGSDR_EXP=$e -c gpu_sdr_b200/csrc/pfb_kernels.cu -o build_exp/pfb_$e.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build_exp/libgsdr_exp$e.so build_exp/pfb_$e.o build/obj/chirp_kernels.o build/obj/direct_kernels.o build/obj/tones_kernels.o build/obj/rx.o build/obj/tx.o build/obj/host.o build/obj/hostlogic.o
done
