O=gpurun_out/r2j; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py -q -x --timeout 90 > $O/tests_quick.log 2>&1; echo "quick tests rc=$? $(tail -1 $O/tests_quick.log)"
timeout 300 python tools/kernel_times.py --streams 64 --frames 8 > $O/kt.txt 2>&1; tail -8 $O/kt.txt
B="python bench.py --steps 4 --warmup 2 --skip-e2e --skip-verify --no-cpu-baseline --no-extra --groups 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inter16 -s 2 -c 1 -f -o $O/r02_inter_cpasync2 $B > $O/ncu.log 2>&1; echo "ncu rc=$?"
