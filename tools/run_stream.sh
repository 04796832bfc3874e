#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_deflate.py -x -q -m gpu -k "streaming or public_interface" > gpurun_out/r2_stream_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2_stream_pytest.log | cut -c1-400
