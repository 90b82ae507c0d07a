# round-2 run 4: (a) debug dump of the packed loop filter's 1080p mismatch, (b) the whole GPU
# suite + bench on the default library (round-1 loop filter + all round-2 host-side changes)
O=gpurun_out/r2d; mkdir -p $O
VP8B200_LIB=$PWD/gpurun_variants_packedlf.so timeout 120 python tools/lf_debug.py > $O/lf_debug.txt 2>&1; echo "debug rc=$?"; head -40 $O/lf_debug.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 180 > $O/tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/tests.log)"; grep -E "^FAILED|^ERROR" $O/tests.log | head -20
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/smoke.log)"
C=$(ls streams/c5_1080p_s*.ivf); N=$(nproc)
run() { name=$1; shift; VP8B200_SYNC=block timeout 120 hostdec/_build/b200bench "$@" $C > $O/e2e_$name.json 2>$O/e2e_$name.err; echo "$name: $(cat $O/e2e_$name.json | cut -c1-400)"; }
run tN_pipe --threads $N --streams 64 --repeat 4 --touch --pipeline
run t64_block --threads 64 --streams 64 --repeat 4 --touch
run t2N_pipe --threads $((2*N)) --streams 64 --repeat 4 --touch --pipeline
run s1_delay --threads 1 --streams 1 --repeat 4 --touch --delay
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -3 $O/bench.err
