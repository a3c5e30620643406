import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import sfo
from starflood_b200 import engine as e
e.init()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
pos, vel, mas, rad = sfo.initcond(N, outflow=0.0, sph=True)
sim = e.Sim(N, sph=True); sim.pos[:] = pos; sim.mas[:] = mas; sim.rad[:] = rad
for r in range(2):
    sim.acc[:] = 0
    e.solve_sph(sim); t = e.timings(); print('SPH N', N, 'solve_ms', t.solve_ms, flush=True)
