O=gpurun_out/r2ab; mkdir -p $O
for v in base spin spin16; do
  if [ $v = base ]; then unset VP8B200_LIB; else export VP8B200_LIB=$PWD/gpurun_variants_$v.so; fi
  timeout 300 python tools/kernel_times.py --streams 1 --frames 2 --reps 3 > $O/kt1_$v.txt 2>&1; echo "$v 1 stream: $(grep '^frame  0' $O/kt1_$v.txt | cut -c60-110)"
  timeout 300 python tools/kernel_times.py --streams 64 --frames 30 --reps 2 > $O/kt64_$v.txt 2>&1; grep "^frame" $O/kt64_$v.txt | awk '{k+= ($4==0)? $16:0; if ($4==1) s+=$16} END {print "   64 streams: key", k, "P total", s}'
done
