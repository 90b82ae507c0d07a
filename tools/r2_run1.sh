# round-2 run 1: parity suite after the lazy-fetch / set-reference changes + e2e harness sweep
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > $O/gpu.txt; nproc >> $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/tests.log)"
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/smoke.log)"
C=$(ls streams/c5_1080p_s*.ivf)
N=$(nproc)
run() { name=$1; shift; VP8B200_SYNC=block timeout 120 hostdec/_build/b200bench "$@" $C > $O/e2e_$name.json 2>$O/e2e_$name.err; echo "$name: $(cat $O/e2e_$name.json | cut -c1-330)"; }
run t64_block --threads 64 --streams 64 --repeat 4 --touch
run t64_delay --threads 64 --streams 64 --repeat 4 --touch --delay
run tN_pipe --threads $N --streams 64 --repeat 4 --touch --pipeline
run t2N_pipe --threads $((2*N)) --streams 64 --repeat 4 --touch --pipeline
run tN_pipe_delay --threads $N --streams 64 --repeat 4 --touch --pipeline --delay
run tN_pipe_notouch --threads $N --streams 64 --repeat 4 --pipeline
run t2N_pipe_s128 --threads $((2*N)) --streams 128 --repeat 3 --touch --pipeline
VP8B200_NO_DEVICE=1 hostdec/_build/b200bench --threads $N --streams 64 --repeat 4 $C > $O/parse_only_tN.json; echo "parse_only: $(cut -c1-200 $O/parse_only_tN.json)"
run s1_block --threads 1 --streams 1 --repeat 4 --touch
run s1_delay --threads 1 --streams 1 --repeat 4 --touch --delay
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-600 $O/bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -3 $O/bench.err
