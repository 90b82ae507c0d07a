# round-2 run 2: new packed loop filter - parity suite, per-kernel times (variants), single-stream table
O=gpurun_out/r2b; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/tests.log)"
grep -E "FAILED|ERROR" $O/tests.log | head -20
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/smoke.log)"
for v in default w4 w2pf4; do
  if [ $v = default ]; then unset VP8B200_LIB; else export VP8B200_LIB=$PWD/gpurun_variants_$v.so; fi
  timeout 300 python tools/kernel_times.py --streams 64 --frames 8 > $O/kt_$v.txt 2>&1; echo "== $v"; tail -8 $O/kt_$v.txt
  timeout 300 python tools/kernel_times.py --streams 1 --frames 8 > $O/kt1_$v.txt 2>&1; tail -4 $O/kt1_$v.txt
done
unset VP8B200_LIB
timeout 600 python bench.py --steps 20 --warmup 5 --skip-e2e --no-cpu-baseline --no-extra > $O/bench_kernels.json 2> $O/bench_kernels.err; echo "bench rc=$?"; cut -c1-2500 $O/bench_kernels.json; tail -3 $O/bench_kernels.err
