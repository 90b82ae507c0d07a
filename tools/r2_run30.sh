O=gpurun_out/r2ac; mkdir -p $O
for v in iprof nopoll; do
export VP8B200_LIB=$PWD/gpurun_variants_$v.so
timeout 300 python tools/kernel_times.py --streams 1 --frames 2 --reps 3 > $O/kt1_$v.txt 2>&1; echo "== $v"; grep -A2 "^frame  0" $O/kt1_$v.txt | cut -c1-330
done
