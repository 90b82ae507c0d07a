O=gpurun_out/r2aq; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > $O/tests_gpu.log 2>&1; echo "gpu tests rc=$? $(tail -1 $O/tests_gpu.log)"; grep -B5 "Error\|FAILED" $O/tests_gpu.log | tail -30
C=$(ls streams/c5_1080p_s100.ivf)
for mode in "--delay" ""; do
  for co in 1 0; do
    VP8B200_COALESCE=$co hostdec/_build/b200bench --threads 1 --streams 1 --repeat 6 --touch $mode $C > $O/s1_co${co}_d${mode}.json 2>$O/err.txt; python - <<PY
import json
d=json.loads(open("$O/s1_co${co}_d${mode}.json").read().strip().splitlines()[-1])
print("single stream coalesce=$co mode='$mode' fps", d["fps"], "cpu decode", d["cpu_ms_per_frame_decode"], "get", d["cpu_ms_per_frame_get_frame"], "blocked", d["blocked_ms_per_frame"])
PY
  done
done
VP8B200_NO_DEVICE=1 hostdec/_build/b200bench --threads 1 --streams 1 --repeat 6 $C 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('parse only fps', d['fps'])"
