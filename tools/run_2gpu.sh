#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_2gpu.err; cut -c1-300 gpurun_out/r2_bench_2gpu.json
timeout 600 python -m pytest tests/test_gpu_deflate.py -x -q -m gpu -k "pool" 2>&1 | tail -1
