#!/bin/bash
timeout 300 python tools/phase_times.py 256 6 2>&1 | tail -1 | cut -c1-150
timeout 900 python -m pytest tests/test_gpu_deflate.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
