O=gpurun_out/r2af; mkdir -p $O
for v in base pk1 pk2 pk3 ps1 ps3 pk2ps3; do
  if [ $v = base ]; then unset VP8B200_LIB; else export VP8B200_LIB=$PWD/gpurun_variants_$v.so; fi
  timeout 120 python tools/kernel_times.py --streams 1 --frames 2 --reps 3 > $O/kt1_$v.txt 2>&1; echo "$v 1 stream: $(grep '^frame  0' $O/kt1_$v.txt | cut -c60-110)"
done
