FB200_SPARSE=8 timeout 600 ncu --set full --import-source on --clock-control none -k regex:sparse_pool -c 1 -o gpurun_out/prof_sparse_pool python tools/phase_times.py 64 6 > gpurun_out/ncu_sparse_pool.log 2>&1
FB200_SPARSE=6 timeout 600 ncu --set full --import-source on --clock-control none -k regex:sparse_parse -c 1 -o gpurun_out/prof_sparse6 python tools/phase_times.py 64 6 > gpurun_out/ncu_sparse6.log 2>&1
tail -3 gpurun_out/ncu_sparse_pool.log
