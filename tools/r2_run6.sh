# round-2 run 6: TMA k_inter16 parity + timing; coalescer with futex wake / priority
O=gpurun_out/r2f; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py -q -x --timeout 90 > $O/tests_quick.log 2>&1; echo "quick tests rc=$? $(tail -1 $O/tests_quick.log)"; grep -E "^FAILED|^ERROR|Error" $O/tests_quick.log | head
if grep -q "passed" $O/tests_quick.log && ! grep -q "failed" $O/tests_quick.log; then
  timeout 600 python -m pytest tests -m gpu -q --timeout 180 > $O/tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/tests.log)"; grep -E "^FAILED|^ERROR" $O/tests.log | head -20
fi
timeout 300 python tools/kernel_times.py --streams 64 --frames 8 > $O/kt.txt 2>&1; tail -8 $O/kt.txt
timeout 300 python tools/kernel_times.py --streams 1 --frames 8 > $O/kt1.txt 2>&1; tail -3 $O/kt1.txt
C=$(ls streams/c5_1080p_s*.ivf); N=$(nproc)
run() { name=$1; shift; VP8B200_SYNC=block timeout 120 hostdec/_build/b200bench "$@" $C > $O/e2e_$name.json 2>$O/e2e_$name.err; echo "$name: $(python -c "
import json,sys
d=json.load(open('$O/e2e_$name.json')); print({k:d[k] for k in ('fps','threads','kernel_launches','engine_batches','engine_frames','cpu_ms_per_frame_decode','cpu_ms_per_frame_get_frame','blocked_ms_per_frame')})" 2>&1 | tail -1)"; }
run co_t2N_pipe --threads $((2*N)) --streams 64 --repeat 4 --touch --pipeline
VP8B200_BATCH_WINDOW_US=0 run co_w0_t2N --threads $((2*N)) --streams 64 --repeat 4 --touch --pipeline
VP8B200_BATCH_WINDOW_US=3000 run co_w3000_t2N --threads $((2*N)) --streams 64 --repeat 4 --touch --pipeline
VP8B200_COALESCE=0 run direct_t2N_pipe --threads $((2*N)) --streams 64 --repeat 4 --touch --pipeline
run co_t3N_pipe --threads $((3*N)) --streams 64 --repeat 4 --touch --pipeline
run co_s1_delay --threads 1 --streams 1 --repeat 4 --touch --delay
B="python bench.py --steps 30 --warmup 6 --skip-e2e --no-cpu-baseline --no-extra"
timeout 300 $B > $O/bench.json 2> $O/bench.err; echo "== bench rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "lf ms", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], {k:v.get("ms_total") for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)})
except Exception as e: print("no result", e, open("$O/bench.err").read()[-300:])
PY
