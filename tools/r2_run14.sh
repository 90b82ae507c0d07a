O=gpurun_out/r2m; mkdir -p $O
B2="python bench.py --steps 4 --warmup 2 --skip-e2e --skip-verify --no-cpu-baseline --no-extra --groups 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_loopfilter -s 3 -c 1 -f -o $O/r02_lf $B2 > $O/ncu_lf.log 2>&1; echo "ncu lf rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 6 --warmup 3 --skip-e2e --skip-verify --no-cpu-baseline --no-extra > $O/launch.log 2>&1; echo "ncu launches rc=$?"
python bench.py --steps 30 --warmup 6 --skip-e2e --no-cpu-baseline --no-extra > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.json
