#!/bin/sh
# Build a kernel-experiment variant of libgsdr.so: tools/build_variant.sh NAME "-DGSDR_WP_LA=2 ..."
# Output: variants/libgsdr_NAME.so (git-ignored; select with GSDR_LIB_PATH).  Needs `make` to have run.
set -e
cd "$(dirname "$0")/.."
mkdir -p variants build/var
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O2,-ffp-contract=off,-fno-fast-math $2 \
     -c gpu_sdr_b200/csrc/pfb_kernels.cu -o build/var/pfb_$1.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/libgsdr_$1.so build/var/pfb_$1.o \
     build/obj/chirp_kernels.o build/obj/direct_kernels.o build/obj/tones_kernels.o build/obj/rx.o build/obj/tx.o build/obj/host.o build/obj/hostlogic.o
