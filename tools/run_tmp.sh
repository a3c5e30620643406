timeout 1500 python -m pytest tests/test_gpu_peer.py -x -q -m gpu 2>&1 | tail -5
nvidia-smi -L | head -3
