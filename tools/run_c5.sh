#!/bin/bash
for k in 1 2 3; do timeout 900 python -m pytest tests/test_gpu_deflate.py -x -q -m gpu 2>&1 | tail -25 | grep -v "^$" | tail -12; done
