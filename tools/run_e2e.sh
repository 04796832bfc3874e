for s in 8,32 4,16 8,64 16,64 32,128 2,8 8,24; do FB200_SLAB=$s timeout 120 python tools/e2e_times.py 256 6 2>&1 | tail -1; done
timeout 120 python tools/phase_times.py 256 6 2>&1 | tail -1 | cut -c1-120
