timeout 3000 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --no-c5 --no-vis --no-sph --no-cpu > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; echo rc=$?; tail -3 gpurun_out/r02_bench_n1_b.err
