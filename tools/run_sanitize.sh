for t in memcheck racecheck; do
timeout 1500 compute-sanitizer --tool $t --print-limit 5 python tools/sanitize_small.py > gpurun_out/r2_sanitize_$t.log 2>&1
echo "== $t: $(grep -c 'Invalid\|Race reported\|hazard' gpurun_out/r2_sanitize_$t.log) findings"; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|sanitize run ok\|sparse repairs" gpurun_out/r2_sanitize_$t.log | tail -3
done
