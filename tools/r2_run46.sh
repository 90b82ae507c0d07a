O=gpurun_out/r2as; mkdir -p $O
export VP8B200_LIB=$PWD/gpurun_variants_shfma.so
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py -q -x --timeout 90 > $O/tests_quick.log 2>&1; echo "shfma quick tests rc=$? $(tail -1 $O/tests_quick.log)"
for v in base shfma base shfma; do
  if [ $v = base ]; then unset VP8B200_LIB; else export VP8B200_LIB=$PWD/gpurun_variants_$v.so; fi
  timeout 300 python tools/kernel_times.py --streams 64 --frames 8 > $O/kt_$v.txt 2>&1
  echo "$v: $(grep 'type 1' $O/kt_$v.txt | awk '{s+=$14; n++} END {printf "inter avg %.4f ms over %d P frames", s/n, n}')"
done
B="python bench.py --steps 30 --warmup 6 --skip-e2e --no-cpu-baseline --no-extra"
for v in base shfma base shfma; do
  if [ $v = base ]; then unset VP8B200_LIB; else export VP8B200_LIB=$PWD/gpurun_variants_$v.so; fi
  timeout 300 $B > $O/bench_$v.json 2> $O/bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$v.json").read().strip().splitlines()[-1])
    print("$v value", d["value"], "ms/step", d["ms_per_step"], {k:v.get("ms_total") for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)})
except Exception as e: print("no result", e, open("$O/bench_$v.err").read()[-300:])
PY
done
