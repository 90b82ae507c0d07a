# round-end measurement set (run under gpurun): GPU tests, smoke, both bench arms, ncu full captures of
# the three big kernels and the launch list of a bench run.  Evidence is copied to profiles/ by hand.
O=gpurun_out/final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > $O/tests_gpu.log 2>&1; echo "gpu tests rc=$? $(tail -1 $O/tests_gpu.log)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/smoke.log)"
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?"; tail -c 1500 $O/bench_default.json
timeout 900 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?"; tail -c 600 $O/bench_reference.json
B2="python bench.py --steps 4 --warmup 2 --skip-e2e --skip-verify --no-cpu-baseline --no-extra --groups 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inter16 -s 2 -c 1 -f -o $O/r02_inter $B2 > $O/ncu_inter.log 2>&1; echo "ncu inter rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_loopfilter -s 3 -c 1 -f -o $O/r02_lf $B2 > $O/ncu_lf.log 2>&1; echo "ncu lf rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intra -s 0 -c 4 -f -o $O/r02_intra python bench.py --steps 2 --warmup 0 --skip-e2e --skip-verify --no-cpu-baseline --no-extra --groups 1 --stagger 0 > $O/ncu_intra.log 2>&1; echo "ncu intra rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 6 --warmup 3 --skip-e2e --skip-verify --no-cpu-baseline --no-extra > $O/launch.log 2>&1; echo "ncu launches rc=$?"
