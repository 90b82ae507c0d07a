O=gpurun_out/r2ae; mkdir -p $O
export VP8B200_LIB=$PWD/gpurun_variants_iprof.so
timeout 300 python tools/kernel_times.py --streams 1 --frames 2 --reps 3 > $O/kt1.txt 2>&1; grep -A9 "^frame  0" $O/kt1.txt | cut -c1-330
