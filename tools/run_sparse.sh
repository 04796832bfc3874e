timeout 120 python tools/phase_times.py 256 6 2>&1 | tail -1 | sed -e 's/hash_link=[0-9.]* //' | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_deflate.py -x -q -k "token or sparse or sharded or bit_exact" 2>&1 | tail -2
