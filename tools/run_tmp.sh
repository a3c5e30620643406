timeout 1500 python -m pytest tests/test_gpu_peer.py tests/test_gpu_vis.py tests/test_gpu_resident.py -x -q -m gpu 2>&1 | tail -15
