O=gpurun_out/r2ax; mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/smoke.log)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py tests/test_gpu_hostdec.py tests/test_gpu_fullsize.py -q -x --timeout 200 > $O/tests_quick.log 2>&1; echo "quick tests rc=$? $(tail -1 $O/tests_quick.log)"
for v in base head base head; do
  if [ $v = base ]; then unset VP8B200_LIB; else export VP8B200_LIB=$PWD/gpurun_variants_$v.so; fi
  timeout 300 python tools/kernel_times.py --streams 64 --frames 30 --reps 2 > $O/kt64_$v.txt 2>&1; grep "^frame" $O/kt64_$v.txt | awk '{k+= ($4==0)? $16:0; if ($4==1) {s+=$16}} END {print "'$v' 64 streams: intra key", k, "P total", s}'
done
