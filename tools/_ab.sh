python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
for v in b2 b3 s128 b2s64; do
  echo "== $v"
  VP8B200_LIB=$PWD/gpurun_variants_$v.so timeout 60 python tools/kernel_times.py --streams 1 --frames 8 2>&1 | grep -E "frame  0"
  VP8B200_LIB=$PWD/gpurun_variants_$v.so timeout 60 python tools/kernel_times.py --streams 64 --frames 8 2>&1 | grep -E "frame  0|frame  1|frame  7"
done
