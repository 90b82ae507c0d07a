mkdir -p gpurun_out/v8
python -m pytest tests -m gpu -q > gpurun_out/v8/tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/v8/tests.log)"
python __graft_entry__.py smoke > gpurun_out/v8/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/v8/smoke.log)"
B="python bench.py --skip-e2e --skip-verify --no-cpu-baseline --steps 60 --warmup 5"
for cfg in "1 0" "2 1" "4 1" "4 0" "3 1"; do
  set -- $cfg
  $B --groups $1 --stagger $2 > gpurun_out/v8/g$1_s$2.json 2> gpurun_out/v8/g$1_s$2.err
  python -c "import json;d=json.load(open('gpurun_out/v8/g$1_s$2.json'));print('groups $1 stagger $2:',d['value'],d['ms_per_step'],d['roofline']['achieved'])"
done
python tools/kernel_times.py --streams 1 --frames 8 2>&1 | tail -8
python tools/kernel_times.py --streams 64 --frames 8 2>&1 | tail -8
