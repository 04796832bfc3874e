# compute-sanitizer over tools/sanitize_small.py: memcheck and racecheck, logs into gpurun_out/ (copied to profiles/ by hand)
for t in ${SANITIZE_TOOLS:-memcheck racecheck}; do
timeout 1500 compute-sanitizer --tool $t --print-limit 200 python tools/sanitize_small.py > gpurun_out/r2_sanitize_$t.log 2>&1
echo "== $t: $(grep -c 'Invalid\|Race reported' gpurun_out/r2_sanitize_$t.log) findings"; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|sanitize run ok\|sparse repairs" gpurun_out/r2_sanitize_$t.log | tail -3
done
# one line per site pair
grep -h "Race reported\|     and " gpurun_out/r2_sanitize_racecheck.log 2>/dev/null | sed 's/=========//; s/+0x[0-9a-f]*//g; s/(.*)//' | sort | uniq -c | sort -rn | head -60 > gpurun_out/r2_racecheck_sites.txt
