FB200_SPARSE=wide,4,2,4,4 timeout 600 ncu --set full --import-source on --clock-control none -k regex:sparse_parse -c 1 -o gpurun_out/prof_sparse1 python tools/phase_times.py 64 6 > gpurun_out/ncu_sparse1.log 2>&1
tail -3 gpurun_out/ncu_sparse1.log
