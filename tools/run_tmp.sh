for parts in 8 4 2 1; do for pf in 0 1; do SFE_WALK_PF=$pf timeout 120 python tools/walk_slice_once.py $parts 2>&1 | tail -2; done; done
