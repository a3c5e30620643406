"""Walk of a fraction of the targets (what one rank of a multi-GPU run does): python tools/walk_slice_once.py [parts]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from starflood_b200 import engine as e, ics
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = 1 << 20
e.init()
pos, vel, mas = ics.uniform_ball(N)
sim = e.Sim(N, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas
e.tree_init(sim.tree, N + 64); sim.tree.bounds[:] = e.root_box(sim)
assert e.tree_build(sim.tree, N, sim.pos, sim.mas) == 0
for fl in (e.WALK_EXACT, e.WALK_FAST):
    for r in range(3):
        rc = e.solve_gravity_part_tree(sim, fl, i0=(N // parts) * 3 % N if parts > 3 else 0, ni=N // parts); t = e.timings()
    print('walk of N/%d targets (index slice), flavour %d: %.3f ms  PF=%s' % (parts, fl, t.solve_ms, os.environ.get('SFE_WALK_PF', 'auto')))
