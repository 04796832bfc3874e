#!/bin/bash
mkdir -p gpurun_out
for p in 32768 65536; do FB200_STREAM_PART=$p timeout 300 python tools/stream_times.py 256 4 2>&1 | tail -1; done
timeout 1200 python -m pytest tests/test_gpu_deflate.py tests/test_gpu_inflate.py -x -q -m gpu > gpurun_out/r2_stream_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_stream_pytest.log | cut -c1-300
