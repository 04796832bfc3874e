timeout 600 ncu --set full --import-source on --clock-control none -k regex:inflate_members -s 2 -c 1 -o gpurun_out/prof_inf6 python tools/inflate_times.py 64 > gpurun_out/ncu_inf6.log 2>&1
tail -2 gpurun_out/ncu_inf6.log
