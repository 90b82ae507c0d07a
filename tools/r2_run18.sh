O=gpurun_out/r2q; mkdir -p $O
timeout 600 python tools/kernel_times.py --streams 64 --frames 30 > $O/kt30.txt 2>&1; grep "^frame" $O/kt30.txt
timeout 600 python tools/kernel_times.py --streams 1 --frames 30 > $O/kt30_s1.txt 2>&1; grep "^frame" $O/kt30_s1.txt | head -12
