#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k streaming 2>&1 | tail -8
