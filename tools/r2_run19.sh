O=gpurun_out/r2r; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py tests/test_gpu_hostdec.py -q -x --timeout 120 > $O/tests_quick.log 2>&1; echo "quick tests rc=$? $(tail -1 $O/tests_quick.log)"
for v in base oldintra base oldintra; do
  if [ $v = base ]; then unset VP8B200_LIB; else export VP8B200_LIB=$PWD/gpurun_variants_$v.so; fi
  timeout 300 python tools/kernel_times.py --streams 64 --frames 30 > $O/kt_$v.txt 2>&1
  echo "$v 64: key $(grep 'frame  0' $O/kt_$v.txt | awk '{print $14}') ms; P-frame intra total $(grep 'type 1' $O/kt_$v.txt | awk '{s+=$14} END {printf "%.3f", s}') ms"
  timeout 300 python tools/kernel_times.py --streams 1 --frames 8 > $O/kt1_$v.txt 2>&1
  echo "$v 1: key $(grep 'frame  0' $O/kt1_$v.txt | awk '{print $14}') ms"
done
