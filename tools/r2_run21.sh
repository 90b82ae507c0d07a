O=gpurun_out/r2t; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py tests/test_gpu_hostdec.py -q -x --timeout 120 > $O/tests_quick.log 2>&1; echo "quick tests rc=$? $(tail -1 $O/tests_quick.log)"
timeout 300 python tools/kernel_times.py --streams 1 --frames 2 --reps 3 > $O/kt1b.txt 2>&1; grep "^frame  0" $O/kt1b.txt | cut -c1-160
timeout 300 python tools/kernel_times.py --streams 64 --frames 30 --reps 2 > $O/kt64b.txt 2>&1; grep "^frame" $O/kt64b.txt | awk '{k+= ($4==0)? $16:0; if ($4==1) s+=$16} END {print "64 streams: key", k, "P total", s}'
export VP8B200_LIB=$PWD/gpurun_variants_iprof.so
timeout 300 python tools/kernel_times.py --streams 1 --frames 2 --reps 3 > $O/kt1.txt 2>&1; grep -A1 "^frame  0" $O/kt1.txt | cut -c1-160
timeout 300 python tools/kernel_times.py --streams 64 --frames 2 --reps 3 > $O/kt64.txt 2>&1; grep -A1 "^frame  0" $O/kt64.txt | cut -c1-160
