O=gpurun_out/r2au; mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/smoke.log)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py tests/test_gpu_hostdec.py tests/test_gpu_fullsize.py -q -x --timeout 200 > $O/tests_quick.log 2>&1; echo "quick tests rc=$? $(tail -1 $O/tests_quick.log)"
for v in base head base head; do
  if [ $v = base ]; then unset VP8B200_LIB; else export VP8B200_LIB=$PWD/gpurun_variants_$v.so; fi
  timeout 300 python tools/kernel_times.py --streams 1 --frames 2 --reps 3 > $O/kt1_$v.txt 2>&1; echo "$v 1 stream: $(grep '^frame  0' $O/kt1_$v.txt | cut -c60-110)"
  timeout 300 python tools/kernel_times.py --streams 64 --frames 30 --reps 2 > $O/kt64_$v.txt 2>&1; grep "^frame" $O/kt64_$v.txt | awk '{k+= ($4==0)? $16:0; if ($4==1) {s+=$16; i+=$14; l+=$18}} END {print "   64 streams: intra key", k, "P total", s, "| inter total", i, "| lf total", l}'
done
B="python bench.py --steps 30 --warmup 6 --skip-e2e --no-cpu-baseline --no-extra"
for v in base head base head; do
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
