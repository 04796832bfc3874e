#!/bin/bash
# round-2 final pass on one GPU: all parity tests, the bench line, the reference arm, the ncu launch list of the bench command,
# and --set full captures of the kernels that changed since the first pass
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_final.log
timeout 1200 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-c5 --skip-stream > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu rc=$?"
NCU="timeout 900 ncu --set full --import-source on --clock-control none"
if [ -n "$FULL_CAPTURES" ]; then
$NCU -k regex:pack_blocks -s 2 -c 1 -f -o gpurun_out/r2_pack_huff4g python tools/phase_times.py 4096 1 > gpurun_out/r2_ncu_pack_huff4g.log 2>&1; tail -1 gpurun_out/r2_ncu_pack_huff4g.log
$NCU -k regex:bit_counts_eager -s 2 -c 1 -f -o gpurun_out/r2_bit_counts_eager python tools/phase_times.py 4096 1 > gpurun_out/r2_ncu_eager.log 2>&1; tail -1 gpurun_out/r2_ncu_eager.log
fi
$NCU -k regex:hash_link -s 2 -c 1 -f -o gpurun_out/r2_hash_link_c python tools/phase_times.py 256 6 > gpurun_out/r2_ncu_hash_link_c.log 2>&1; tail -1 gpurun_out/r2_ncu_hash_link_c.log
