# round-end measurement set (run under gpurun): GPU tests, smoke, e2e CPU split, both bench arms,
# ncu launch list + one full capture of k_loopfilter, per-config table. Outputs: gpurun_out/v8/ ->
# the ones that are evidence are copied to profiles/ by hand.
mkdir -p gpurun_out/v8
python -m pytest tests -m gpu -q > gpurun_out/v8/tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/v8/tests.log)"
python __graft_entry__.py smoke > gpurun_out/v8/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/v8/smoke.log)"
C=$(ls streams/c5_1080p_s*.ivf)
for mode in new ref; do
  if [ $mode = ref ]; then export VP8B200_TOKENS=ref; else unset VP8B200_TOKENS; fi
  VP8B200_SYNC=block hostdec/_build/b200bench --threads 64 --streams 64 --repeat 4 $C > gpurun_out/v8/e2e_block_$mode.json
  VP8B200_NO_DEVICE=1 hostdec/_build/b200bench --threads 64 --streams 64 --repeat 4 $C > gpurun_out/v8/parse_only_$mode.json
done
unset VP8B200_TOKENS
VP8B200_SYNC=block hostdec/_build/b200bench --threads 32 --streams 64 --repeat 4 $C > gpurun_out/v8/e2e_block_new_t32.json
VP8B200_SYNC=block hostdec/_build/b200bench --threads 128 --streams 128 --repeat 3 $C > gpurun_out/v8/e2e_block_new_s128.json
python bench.py --impl reference > gpurun_out/v8/bench_ref.json 2> gpurun_out/v8/bench_ref.err
python bench.py > gpurun_out/v8/bench.json 2> gpurun_out/v8/bench.err
B="python bench.py --steps 4 --warmup 2 --skip-e2e --skip-verify --no-cpu-baseline --groups 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/v8/r01_launches_v8.csv $B > gpurun_out/v8/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_loopfilter -s 3 -c 1 -f -o gpurun_out/v8/r01_lf_v8 $B > gpurun_out/v8/ncu2.log 2>&1
for f in gpurun_out/v8/*.json; do echo "== $f"; cat $f; echo; done
python tools/config_table.py --out gpurun_out/v8/r01_configs.json > gpurun_out/v8/cfg.log 2>&1; tail -6 gpurun_out/v8/cfg.log | cut -c1-400
