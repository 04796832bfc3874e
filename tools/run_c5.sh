#!/bin/bash
timeout 300 python tools/phase_times.py 4096 1 2>&1 | tail -1
FB200_BITCOUNTS=lanes timeout 300 python tools/phase_times.py 4096 1 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_deflate.py tests/test_gpu_fullsize.py -x -q -m gpu -k "huffman or simple or block_range or fullsize or compress_bit_exact or streaming or kats or golden" 2>&1 | tail -3
