#!/bin/bash
# round-2 GPU pass: parity tests, the bench line, and one ncu --set full capture per config's dominant kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest.log
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_bench.json
NCU="timeout 900 ncu --set full --import-source on --clock-control none"
$NCU -k regex:sparse_parse -s 2 -c 1 -f -o gpurun_out/r2_sparse_L6 python tools/phase_times.py 256 6 > gpurun_out/r2_ncu_sparse_L6.log 2>&1; tail -1 gpurun_out/r2_ncu_sparse_L6.log
$NCU -k regex:sparse_parse -s 2 -c 1 -f -o gpurun_out/r2_sparse_L9 python tools/phase_times.py 256 9 > gpurun_out/r2_ncu_sparse_L9.log 2>&1; tail -1 gpurun_out/r2_ncu_sparse_L9.log
$NCU -k regex:inflate_members_par -s 2 -c 1 -f -o gpurun_out/r2_inflate_1024 python tools/inflate_times.py 1024 > gpurun_out/r2_ncu_inflate_1024.log 2>&1; tail -1 gpurun_out/r2_ncu_inflate_1024.log
$NCU -k regex:build_blocks -s 2 -c 1 -f -o gpurun_out/r2_build_huff4g python tools/phase_times.py 4096 1 > gpurun_out/r2_ncu_build_huff4g.log 2>&1; tail -1 gpurun_out/r2_ncu_build_huff4g.log
$NCU -k regex:hash_link -s 2 -c 1 -f -o gpurun_out/r2_hash_link python tools/phase_times.py 256 6 > gpurun_out/r2_ncu_hash_link.log 2>&1; tail -1 gpurun_out/r2_ncu_hash_link.log
ls -la gpurun_out/r2_*.ncu-rep
