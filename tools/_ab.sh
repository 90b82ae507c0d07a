mkdir -p gpurun_out/ab2
for v in base nofence early early_nofence; do
  echo "== $v"
  VP8B200_LIB=$PWD/gpurun_variants_$v.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
  VP8B200_LIB=$PWD/gpurun_variants_$v.so python tools/kernel_times.py --streams 1 --frames 4 2>&1 | tail -2
  VP8B200_LIB=$PWD/gpurun_variants_$v.so python tools/kernel_times.py --streams 64 2>&1 | tail -3
  VP8B200_LIB=$PWD/gpurun_variants_$v.so python bench.py --skip-e2e --skip-verify --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/ab2/$v.json 2>gpurun_out/ab2/$v.err
  python -c "import json;d=json.load(open('gpurun_out/ab2/$v.json'));print(d['value'],d['ms_per_step'],d['roofline']['achieved'])"
done
