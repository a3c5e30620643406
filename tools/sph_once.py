"""solve_sph() replacement on a C3-like snapshot: python tools/sph_once.py [N]; prints the solve time and the pairs the
cell-list passes evaluated (SFE_SPH_CELL_DIV / SFE_SPH_TPC / SFE_SPH_CULL select the tuning)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from starflood_b200 import engine as e, ics
e.init()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
pos, vel, mas = ics.uniform_ball(N, outflow=0.0)
sim = e.Sim(N, sph=True, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas; sim.rad[:] = 0.1
pairs = C.c_uint64()
for r in range(3):
    sim.acc[:] = 0
    rc = e.solve_sph(sim); t = e.timings()
    e.lib().sfe_sph_pairs_evaluated(C.byref(pairs), None)
    print('SPH N', N, rc, 'solve_ms %.3f' % t.solve_ms, 'pairs evaluated per pass and target %.0f' % (pairs.value / 2.0 / N),
          'tuning', os.environ.get('SFE_SPH_CELL_DIV', '-'), os.environ.get('SFE_SPH_TPC', '-'), os.environ.get('SFE_SPH_CULL', '-'), flush=True)
