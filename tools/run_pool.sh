#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_deflate.py -x -q -m gpu -k "pool or tiny_members or members_batch" 2>&1 | tail -4
