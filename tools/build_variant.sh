#!/bin/sh
# build an A/B variant of libvp8b200.so: tools/build_variant.sh <name> [-DFLAG=...]...
# -> gpurun_variants_<name>.so (git-ignored; select with VP8B200_LIB=$PWD/gpurun_variants_<name>.so)
set -e
name=$1; shift
cd "$(dirname "$0")/../libvpx.opencl_b200"
out=_obj/var_$name; rm -rf $out; mkdir -p $out
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../include -Icsrc --cudart static"
for f in runtime kernels_recon kernels_intra kernels_lf kernels_border; do
  $NV "$@" -c csrc/$f.cu -o $out/$f.o &
done
wait; for f in runtime kernels_recon kernels_intra kernels_lf kernels_border; do test -f $out/$f.o || { echo "compile failed: $f"; exit 1; }; done
$NV -shared -o ../gpurun_variants_$name.so $out/*.o
echo built gpurun_variants_$name.so
