O=gpurun_out/r2h; mkdir -p $O
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc $(tail -1 $O/smoke.log)"
if [ $rc -ne 0 ]; then timeout 200 compute-sanitizer --tool memcheck --print-limit 3 python __graft_entry__.py smoke > $O/sanitizer.log 2>&1; grep -v "Host Frame" $O/sanitizer.log | head -20; exit 0; fi
bash tools/r2_run6.sh
