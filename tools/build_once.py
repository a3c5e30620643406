"""Times tree builds only (no walk): python tools/build_once.py N reps.  SFE_DBG_SCATTER_SEQ="127,0,16" runs one build per mask."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import sfo
from starflood_b200 import engine as e
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
e.init()
pos, vel, mas = sfo.initcond(N)
sim = e.Sim(N, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas
e.tree_init(sim.tree, N + 64); sim.tree.bounds[:] = sfo.root_box(pos, 0)
seq = os.environ.get("SFE_DBG_SCATTER_SEQ")
masks = [int(x) for x in seq.split(",")] if seq else [None] * reps
for m in masks:
    if m is not None: os.environ["SFE_DBG_SCATTER"] = str(m)
    rc = e.tree_build(sim.tree, N, sim.pos, sim.mas); tb = e.timings()
    print('N', N, 'mask', m, rc, 'build_ms', round(tb.build_ms, 3), 'launches', tb.launches)
