import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import sfo
from starflood_b200 import engine as e
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
e.init()
pos, vel, mas = sfo.initcond(N)
sim = e.Sim(N, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas
e.tree_init(sim.tree, N + 64); sim.tree.bounds[:] = sfo.root_box(pos, 0)
for r in range(reps):
    rc = e.tree_build(sim.tree, N, sim.pos, sim.mas); tb = e.timings()
    rc2 = e.solve_gravity_part_tree(sim, fl); tw = e.timings()
    print('N', N, rc, rc2, 'build_ms', round(tb.build_ms, 3), 'h2d', round(tb.h2d_ms, 3), 'walk_ms', round(tw.solve_ms, 3), 'd2h', round(tw.d2h_ms, 3),
          'launches', tb.launches, tw.launches, 'particle-steps/s %.3e' % (N / ((tb.build_ms + tw.solve_ms) * 1e-3)))
