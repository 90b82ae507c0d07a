# round-2 run 3: fixed packed loop filter: quick parity, A/B against the round-1 kernel, ncu counters
O=gpurun_out/r2c; mkdir -p $O
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/smoke.log)"
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py -q -x --timeout 60 > $O/tests_quick.log 2>&1; echo "quick tests rc=$? $(tail -1 $O/tests_quick.log)"
grep -E "^FAILED|^ERROR" $O/tests_quick.log | head
if grep -q "passed" $O/tests_quick.log && ! grep -q "failed" $O/tests_quick.log; then
  timeout 600 python -m pytest tests -m gpu -q --timeout 120 > $O/tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/tests.log)"; grep -E "^FAILED|^ERROR" $O/tests.log | head
fi
B="python bench.py --steps 30 --warmup 6 --skip-e2e --no-cpu-baseline --no-extra"
for v in default w4 oldlf; do
  if [ $v = default ]; then unset VP8B200_LIB; else export VP8B200_LIB=$PWD/gpurun_variants_$v.so; fi
  timeout 300 $B > $O/bench_$v.json 2> $O/bench_$v.err; echo "== bench $v rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$v.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "lf ms", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], {k:v.get("ms_total") for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)})
except Exception as e: print("no result", e, open("$O/bench_$v.err").read()[-300:])
PY
done
unset VP8B200_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_loopfilter -s 3 -c 1 -f -o $O/r02_lf_packed $B --steps 4 --warmup 2 --skip-verify --groups 1 > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu.log
