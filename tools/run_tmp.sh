timeout 900 python -m pytest tests/test_gpu_walk.py -x -q -m gpu -k "solv_tree or golden" 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err; echo rc=$?; tail -3 gpurun_out/r02_bench_n1_a.err
