timeout 600 python -m pytest tests/test_gpu_sph.py -x -q -m gpu 2>&1 | tail -3
for cfg in "4 32 0" "4 32 1" "5 16 1" "6 12 1" "6 8 1" "8 8 1" "8 4 1" "8 4 0"; do
  set -- $cfg
  SFE_SPH_CELL_DIV=$1 SFE_SPH_TPC=$2 SFE_SPH_CULL=$3 timeout 120 python tools/sph_once.py 2>&1 | tail -1
done
