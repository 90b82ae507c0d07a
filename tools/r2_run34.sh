O=gpurun_out/r2ai; mkdir -p $O
for v in iprof; do
  export VP8B200_LIB=$PWD/gpurun_variants_$v.so
  timeout 120 python tools/kernel_times.py --streams 1 --frames 2 --reps 3 > $O/kt1_$v.txt 2>&1; echo "== $v"; grep -A12 '^frame  0' $O/kt1_$v.txt | cut -c1-300
done
