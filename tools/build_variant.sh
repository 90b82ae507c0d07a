#!/bin/sh
# A/B build of libvp8b200 with extra nvcc flags (e.g. -DI16_MINB=7): tools/build_variant.sh NAME [flags...]
# -> gpurun_variants_NAME.so, selected at run time with VP8B200_LIB=$PWD/gpurun_variants_NAME.so
set -e
name=$1; shift
cd "$(dirname "$0")/../libvpx.opencl_b200"
out=_obj/var_$name; rm -rf $out; mkdir -p $out
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../include -Icsrc --cudart static"
for f in runtime kernels_recon kernels_intra kernels_border kernels_lf; do $NV "$@" -c csrc/$f.cu -o $out/$f.o & done
wait
$NV -shared -o ../gpurun_variants_$name.so $out/*.o
echo built gpurun_variants_$name.so
