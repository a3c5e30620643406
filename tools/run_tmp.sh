timeout 120 python tools/bh_once.py 1048576 3 2>&1 | tail -2
timeout 300 python tools/c5_once.py 5 2>&1 | tail -5
