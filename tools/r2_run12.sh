O=gpurun_out/r2k; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py -q -x --timeout 90 > $O/tests_quick.log 2>&1; echo "quick tests rc=$? $(tail -1 $O/tests_quick.log)"
timeout 300 python tools/kernel_times.py --streams 64 --frames 8 > $O/kt.txt 2>&1; tail -8 $O/kt.txt
timeout 300 python tools/kernel_times.py --streams 1 --frames 8 > $O/kt1.txt 2>&1; tail -3 $O/kt1.txt
B="python bench.py --steps 30 --warmup 6 --skip-e2e --no-cpu-baseline --no-extra"
timeout 300 $B > $O/bench.json 2> $O/bench.err; echo "== bench rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "lf ms", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], {k:v.get("ms_total") for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)})
except Exception as e: print("no result", e, open("$O/bench.err").read()[-300:])
PY
B2="python bench.py --steps 4 --warmup 2 --skip-e2e --skip-verify --no-cpu-baseline --no-extra --groups 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inter16 -s 2 -c 1 -f -o $O/r02_inter_direct $B2 > $O/ncu.log 2>&1; echo "ncu rc=$?"
