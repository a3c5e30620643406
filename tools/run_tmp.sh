timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_gpu_walk.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -30
timeout 120 python tools/bh_once.py 1048576 4 2>&1 | tail -3
SFE_TRACE=1 timeout 120 python tools/bh_once.py 1048576 2 2>&1 | tail -9
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_top11_launches.csv python tools/bh_once.py 1048576 2 > /dev/null 2>&1
