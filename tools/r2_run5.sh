# round-2 run 5: coalescer (engine) parity + e2e sweeps; packed loop filter A/B on bench value
O=gpurun_out/r2e; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hostdec.py tests/test_gpu_abi_misc.py -q -x --timeout 120 > $O/tests_quick.log 2>&1; echo "quick tests rc=$? $(tail -1 $O/tests_quick.log)"; grep -E "^FAILED|^ERROR|Error" $O/tests_quick.log | head
timeout 600 python -m pytest tests -m gpu -q --timeout 180 > $O/tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/tests.log)"; grep -E "^FAILED|^ERROR" $O/tests.log | head -20
C=$(ls streams/c5_1080p_s*.ivf); N=$(nproc)
run() { name=$1; shift; VP8B200_SYNC=block timeout 120 hostdec/_build/b200bench "$@" $C > $O/e2e_$name.json 2>$O/e2e_$name.err; echo "$name: $(python -c "
import json,sys
d=json.load(open('$O/e2e_$name.json')); print({k:d[k] for k in ('fps','threads','kernel_launches','engine_batches','engine_frames','cpu_ms_per_frame_decode','cpu_ms_per_frame_get_frame','blocked_ms_per_frame')})" 2>&1 | tail -1)"; }
run co_tN_pipe --threads $N --streams 64 --repeat 4 --touch --pipeline
run co_t2N_pipe --threads $((2*N)) --streams 64 --repeat 4 --touch --pipeline
run co_t64_block --threads 64 --streams 64 --repeat 4 --touch
VP8B200_BATCH_WINDOW_US=1000 run co_w1000_t2N --threads $((2*N)) --streams 64 --repeat 4 --touch --pipeline
VP8B200_BATCH_WINDOW_US=6000 run co_w6000_t2N --threads $((2*N)) --streams 64 --repeat 4 --touch --pipeline
VP8B200_BATCH_WINDOW_US=0 run co_w0_t2N --threads $((2*N)) --streams 64 --repeat 4 --touch --pipeline
VP8B200_COALESCE=0 run direct_t2N_pipe --threads $((2*N)) --streams 64 --repeat 4 --touch --pipeline
run co_s1_delay --threads 1 --streams 1 --repeat 4 --touch --delay
run co_s1_block --threads 1 --streams 1 --repeat 4 --touch
B="python bench.py --steps 30 --warmup 6 --skip-e2e --no-cpu-baseline --no-extra"
for v in default packedlf; do
  if [ $v = default ]; then unset VP8B200_LIB; else export VP8B200_LIB=$PWD/gpurun_variants_$v.so; fi
  timeout 300 $B > $O/bench_$v.json 2> $O/bench_$v.err; echo "== bench $v rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$v.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "lf ms", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], {k:v.get("ms_total") for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)})
except Exception as e: print("no result", e, open("$O/bench_$v.err").read()[-300:])
PY
done
export VP8B200_LIB=$PWD/gpurun_variants_packedlf.so
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py -q --timeout 60 > $O/tests_packed.log 2>&1; echo "packed-lf tests rc=$? $(tail -1 $O/tests_packed.log)"
