timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-pp --no-cpu --no-sph --no-resident --no-vis --no-c5 > gpurun_out/r02_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_walk_exact -f python tools/bh_once.py 1048576 3 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_walk_fast -f python tools/bh_once.py 1048576 3 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cl_pair_kernel --launch-skip 2 --launch-count 2 -o gpurun_out/r02_sph_cl -f python tools/sph_once.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"top_|subtree|xl_|finalize|moments" --launch-skip 11 --launch-count 11 -o gpurun_out/r02_build_kernels -f python tools/bh_once.py 1048576 2 > /dev/null 2>&1
ls -la gpurun_out | tail -8
