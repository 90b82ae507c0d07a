O=gpurun_out/r2ar; mkdir -p $O
C=$(ls streams/c5_1080p_s*.ivf)
run() { name=$1; shift; env "$@" hostdec/_build/b200bench --threads 32 --streams 64 --repeat 4 --touch --pipeline $EXTRA $C > $O/$name.json 2>$O/err.txt; python - <<PY
import json
d=json.loads(open("$O/$name.json").read().strip().splitlines()[-1])
print("$name fps", d["fps"], "decode", d["cpu_ms_per_frame_decode"], "get", d["cpu_ms_per_frame_get_frame"], "blocked", d["blocked_ms_per_frame"], "runq", d["runq_wait_ms_per_frame"], "launches/frame", round(d["kernel_launches"]/d["frames"],3))
PY
}
for rep in 1 2; do
EXTRA="" run std_w1000_$rep VP8B200_BATCH_WINDOW_US=1000
EXTRA="" run std_w500_$rep VP8B200_BATCH_WINDOW_US=500
EXTRA="" run std_w300_$rep VP8B200_BATCH_WINDOW_US=300
EXTRA="--delay" run delay_w1000_$rep VP8B200_BATCH_WINDOW_US=1000
EXTRA="--delay" run delay_w3000_$rep VP8B200_BATCH_WINDOW_US=3000
done
VP8B200_NO_DEVICE=1 hostdec/_build/b200bench --threads 32 --streams 64 --repeat 4 $C 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('parse only fps', d['fps'])"
