for v in gpf r8 base; do
  echo "== $v"
  VP8B200_LIB=$PWD/gpurun_variants_$v.so timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
  VP8B200_LIB=$PWD/gpurun_variants_$v.so timeout 60 python tools/kernel_times.py --streams 1 --frames 4 2>&1 | tail -2
  VP8B200_LIB=$PWD/gpurun_variants_$v.so timeout 60 python tools/kernel_times.py --streams 64 2>&1 | tail -4
done
