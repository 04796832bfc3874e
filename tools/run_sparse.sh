timeout 120 python tools/phase_times.py 256 6 2>&1 | tail -1 | sed -e 's/hash_link=[0-9.]* //' | cut -c1-200
