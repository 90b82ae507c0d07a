O=gpurun_out/r2l; mkdir -p $O
B="python bench.py --steps 30 --warmup 6 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2 3; do
 for v in new r1; do
  if [ $v = r1 ]; then export VP8B200_LIB=$PWD/gpurun_variants_r1inter.so; else unset VP8B200_LIB; fi
  timeout 300 $B > $O/bench_${v}_$rep.json 2> $O/bench_${v}_$rep.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${v}_$rep.json").read().strip().splitlines()[-1])
    print("$v $rep value", d["value"], "ms/step", d["ms_per_step"], {k:v.get("ms_total") for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)}, d.get("clocks"))
except Exception as e: print("no result", e, open("$O/bench_${v}_$rep.err").read()[-300:])
PY
 done
done
