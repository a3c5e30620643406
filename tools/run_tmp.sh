python -m pytest tests/test_gpu_tree.py tests/test_gpu_walk.py tests/test_gpu_configs.py tests/test_gpu_peer.py tests/test_gpu_resident.py -x -q -m gpu 2>&1 | tail -5
python tools/bh_once.py 1048576 3 2>&1 | tail -2
SFE_TRACE=1 python tools/bh_once.py 1048576 2 2>&1 | tail -9
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_top6_launches.csv python tools/bh_once.py 1048576 2 > /dev/null 2>&1
