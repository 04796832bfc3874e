#!/bin/bash
timeout 300 python tools/phase_times.py 256 6 2>&1 | tail -1 | grep -o "L6: [0-9.]* ms/step\|build_blocks=[0-9.]*" | paste - -
FB200_DATA=tar timeout 300 python tools/phase_times.py 256 6 2>&1 | tail -1 | grep -o "L6: [0-9.]* ms/step\|build_blocks=[0-9.]*" | paste - -
timeout 300 python tools/phase_times.py 4096 1 2>&1 | tail -1 | grep -o "L1: [0-9.]* ms/step\|build_blocks=[0-9.]*" | paste - -
timeout 900 python -m pytest tests/test_gpu_deflate.py -x -q -m gpu 2>&1 | tail -2
