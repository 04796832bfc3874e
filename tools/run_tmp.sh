#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_deflate.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python tools/phase_times.py 256 6 2>&1 | grep -o "hash_link=[0-9.]*"
FB200_DATA=tar timeout 120 python tools/phase_times.py 256 6 2>&1 | grep -o "hash_link=[0-9.]*"
