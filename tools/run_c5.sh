#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_b.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_b.log
timeout 300 python tools/phase_times.py 4096 1 2>&1 | tail -1
timeout 300 python tools/phase_times.py 1024 0 2>&1 | tail -1
timeout 300 python tools/phase_times.py 256 6 2>&1 | tail -1
FB200_PROFILE_RANGE=1 timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:inflate_members_par -c 1 -f -o gpurun_out/r2_inflate_1024 python tools/inflate_times.py 1024 > gpurun_out/r2_ncu_inflate_1024.log 2>&1; tail -2 gpurun_out/r2_ncu_inflate_1024.log
