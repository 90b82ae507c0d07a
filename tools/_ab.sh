mkdir -p gpurun_out/ab13
B="python bench.py --skip-e2e --skip-verify --no-cpu-baseline --steps 60 --warmup 5"
for v in cap2 cap3; do
  VP8B200_LIB=$PWD/gpurun_variants_$v.so timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
  for cfg in "2 1" "4 1"; do
    set -- $cfg
    VP8B200_LIB=$PWD/gpurun_variants_$v.so $B --groups $1 --stagger $2 > gpurun_out/ab13/${v}_g$1.json 2> gpurun_out/ab13/${v}_g$1.err
    python -c "import json;d=json.load(open('gpurun_out/ab13/${v}_g$1.json'));print('$v groups $1 stagger $2:',d['value'],d['ms_per_step'],d['roofline']['achieved'])"
  done
done
