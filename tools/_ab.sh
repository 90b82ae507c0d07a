python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -1
python tools/kernel_times.py --streams 1 --frames 8 2>&1 | tail -4
python tools/kernel_times.py --streams 64 --frames 8 2>&1 | tail -5
python bench.py --skip-e2e --skip-verify --no-cpu-baseline --steps 60 --warmup 5 | python -c "import json,sys;d=json.load(sys.stdin);print('bench:',d['value'],d['ms_per_step'],d['roofline']['achieved'],d['roofline']['ms_per_launch'])"
