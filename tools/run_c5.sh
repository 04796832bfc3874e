#!/bin/bash
for k in 1 2; do timeout 300 python tools/stream_times.py 256 4 2>&1 | tail -1; done
FB200_STREAM_TRACE=1 timeout 300 python tools/stream_times.py 256 2 2>&1 | grep "stream part\|trigger" | tail -8
timeout 300 python tools/phase_times.py 4096 1 2>&1 | tail -1 | grep -o "L1: [0-9.]* ms/step\|pack_blocks=[0-9.]*" | paste - -
timeout 600 python -m pytest tests/test_gpu_deflate.py -x -q -m gpu -k "huffman or store or simple or block_range or compress_bit_exact or stored" 2>&1 | tail -2
