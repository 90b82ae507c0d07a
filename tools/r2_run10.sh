O=gpurun_out/r2i; mkdir -p $O
B="python bench.py --steps 4 --warmup 2 --skip-e2e --skip-verify --no-cpu-baseline --no-extra --groups 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inter16 -s 2 -c 1 -f -o $O/r02_inter_cpasync $B > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu.log
