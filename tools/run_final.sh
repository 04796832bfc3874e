#!/bin/bash
# round-2 final pass on one GPU: all parity tests, the bench line, the reference arm, the ncu launch list of the bench command
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_final.log
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-c5 --skip-stream > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu rc=$?"
