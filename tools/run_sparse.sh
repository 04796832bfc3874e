timeout 300 python tools/sparse_check.py > gpurun_out/sparse_repair.log 2>&1; echo "exit $?" >> gpurun_out/sparse_repair.log
cat gpurun_out/sparse_repair.log
(
for t in 0 2; do
FB200_SPARSE=$t timeout 120 python tools/phase_times.py 256 6
done
timeout 120 python tools/phase_times.py 256 9
timeout 120 python tools/phase_times.py 256 4
FB200_PARSE_MODE=1 timeout 120 python tools/phase_times.py 256 9
FB200_PARSE_MODE=1 timeout 120 python tools/phase_times.py 256 4
) 2>&1 > gpurun_out/sparse_times.log
cat gpurun_out/sparse_times.log
