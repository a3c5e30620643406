timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/bh_once.py 1048576 4 2>&1 | tail -3
timeout 300 python tools/c5_once.py 5 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_top14_launches.csv python tools/bh_once.py 1048576 2 > /dev/null 2>&1
