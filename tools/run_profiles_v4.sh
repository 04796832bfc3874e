ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 1 --warmup 1 --bytes 67108864 --skip-inflate --skip-cpu > gpurun_out/ncu_bench_v4.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:sparse_parse -s 2 -c 1 -o gpurun_out/prof_sparse_v4 python tools/phase_times.py 256 6 > gpurun_out/ncu_sparse_v4.log 2>&1
tail -2 gpurun_out/ncu_sparse_v4.log
