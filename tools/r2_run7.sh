O=gpurun_out/r2g; mkdir -p $O
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py smoke > $O/sanitizer.log 2>&1; echo "sanitizer rc=$?"; grep -v "^=========     Host Frame\|^=========         at\|^=========         in " $O/sanitizer.log | head -30
bash tools/r2_run6.sh
