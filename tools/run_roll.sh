#!/bin/bash
mkdir -p gpurun_out
for cfg in 16,1,2 16,1,4 16,1,$((2+256*500)) 16,1,$((3+256*1000)) 32,1,3; do
  FB200_SPARSE_ROLL=$cfg timeout 300 python tools/phase_times.py 256 6 2>&1 | tail -1 | sed "s/^/ROLL=$cfg /" | grep -o "ROLL=[0-9,]*\|ms/step\|sparse_parse=[0-9.]*" | paste - - - 
done
FB200_SPARSE_ROLL=16 timeout 900 ncu --set full --import-source on --clock-control none -k regex:sparse_roll -s 2 -c 1 -f -o gpurun_out/r2_sparse_roll python tools/phase_times.py 256 6 > gpurun_out/r2_ncu_sparse_roll.log 2>&1; tail -1 gpurun_out/r2_ncu_sparse_roll.log
