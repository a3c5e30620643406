python -m pytest tests/test_gpu_tree.py -x -q -m gpu 2>&1 | tail -30
python tools/bh_once.py 1048576 4 2>&1 | tail -3
SFE_TRACE=1 python tools/bh_once.py 1048576 2 2>&1 | tail -9
