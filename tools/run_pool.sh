#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_inflate.py -x -q -m gpu -k "gzip_file" 2>&1 | tail -15
