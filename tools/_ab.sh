# A/B of loop-filter hand-off variants + bench phase stagger (one gpurun call)
mkdir -p gpurun_out/ab3
B="python bench.py --skip-e2e --skip-verify --no-cpu-baseline --steps 60 --warmup 5"
echo "== default (bar, 4 rows/CTA, ring 4): full gpu suite"
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/kernel_times.py --streams 1 --frames 4 2>&1 | tail -3
python tools/kernel_times.py --streams 64 2>&1 | tail -8
for v in spin bar8 bar4r2 bar4pf3; do
  echo "== $v"
  VP8B200_LIB=$PWD/gpurun_variants_$v.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
  VP8B200_LIB=$PWD/gpurun_variants_$v.so python tools/kernel_times.py --streams 64 2>&1 | tail -3
done
for cfg in "4 0" "4 1" "8 1" "2 1"; do
  set -- $cfg
  $B --groups $1 --stagger $2 > gpurun_out/ab3/g$1_s$2.json 2> gpurun_out/ab3/g$1_s$2.err
  python -c "import json;d=json.load(open('gpurun_out/ab3/g$1_s$2.json'));print('groups $1 stagger $2:',d['value'],d['ms_per_step'],d['roofline']['achieved'],d['roofline']['kernels'])"
done
VP8B200_LIB=$PWD/gpurun_variants_spin.so $B --groups 4 --stagger 1 > gpurun_out/ab3/spin_g4_s1.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/ab3/spin_g4_s1.json'));print('spin groups 4 stagger 1:',d['value'],d['ms_per_step'],d['roofline']['achieved'])"
ncu --metrics launch__occupancy_limit_barriers,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,launch__occupancy_limit_warps,sm__warps_active.avg.per_cycle_active,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_loopfilter -s 3 -c 1 --csv --log-file gpurun_out/ab3/lf_occ.csv $B --steps 4 --warmup 2 --groups 1 > /dev/null 2>&1
cat gpurun_out/ab3/lf_occ.csv | tail -9
