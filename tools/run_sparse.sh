(
for v in 0 3 4; do
FB200_SPARSE=$v timeout 120 python tools/phase_times.py 256 6
done
FB200_SPARSE=4 timeout 120 python tools/phase_times.py 256 9
) 2>&1 | sed -e 's/hash_link=.*sparse_parse/sparse_parse/' > gpurun_out/sparse_times.log
cat gpurun_out/sparse_times.log
