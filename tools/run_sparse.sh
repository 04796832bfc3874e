timeout 300 python tools/sparse_check.py > gpurun_out/sparse_wide.log 2>&1; echo "exit $?" >> gpurun_out/sparse_wide.log
tail -n 3 gpurun_out/sparse_wide.log
(
for t in 0 1 2 3 4 0,2,4,4,3 1,2,4,4,3 1,2,4,4,1 1,1,1,1,0 1,3,8,8,0 4,2,4,4,2; do
FB200_SPARSE=$t timeout 120 python tools/phase_times.py 256 6
done
) 2>&1 | sed -e 's/hash_link=.*sparse_parse/sparse_parse/' > gpurun_out/sparse_times.log
cat gpurun_out/sparse_times.log
