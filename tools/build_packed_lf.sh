#!/bin/sh
# A/B build: libvp8b200 with the experimental packed loop filter (csrc/kernels_lf_packed.cu)
# instead of csrc/kernels_lf.cu -> gpurun_variants_packedlf.so (select with VP8B200_LIB=...)
set -e
cd "$(dirname "$0")/../libvpx.opencl_b200"
out=_obj/var_packedlf; rm -rf $out; mkdir -p $out
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../include -Icsrc --cudart static"
for f in runtime kernels_recon kernels_intra kernels_border; do $NV "$@" -c csrc/$f.cu -o $out/$f.o & done
$NV "$@" -c csrc/kernels_lf_packed.cu -o $out/kernels_lf.o
wait
$NV -shared -o ../gpurun_variants_packedlf.so $out/*.o
echo built gpurun_variants_packedlf.so
