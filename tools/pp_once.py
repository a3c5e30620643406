import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import sfo
from starflood_b200 import engine as e
N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
e.init()
pos, vel, mas = sfo.initcond(N)
sim = e.Sim(N, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas
for r in range(reps):
    e.solve_gravity_part_part(sim); t = e.timings()
    print('N', N, 'solve_ms', t.solve_ms, 'TFLOP/s', 20.0 * N * N / (t.solve_ms * 1e-3) / 1e12)
