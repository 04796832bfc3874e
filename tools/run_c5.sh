#!/bin/bash
timeout 300 python tools/inflate_times.py 1 1024 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_inflate.py -x -q -m gpu 2>&1 | tail -2
