"""vis_draw() replacement once at N particles (default 2^20), 1080 x 1080: timings of the host entry point, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from starflood_b200 import engine as e, ics
e.init()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pos, vel, mas = ics.uniform_ball(N)
sim = e.Sim(N, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas
r2 = (pos.reshape(-1, 3).astype(np.float64) ** 2).sum(1)
sim.pot[:] = (-(1.5 - 0.5 * r2) * 4.18879).astype(np.float32)
vis = e.Vis(); e.vis_init(vis, 1080, 1080)
for r in range(reps):
    assert e.vis_draw(vis, sim) == 0, e.last_error()
    t = e.timings()
    print('VIS N', N, 'h2d_ms %.3f kernels_ms %.3f d2h_ms %.3f total_ms %.3f launches %d landed %d' %
          (t.h2d_ms, t.solve_ms, t.d2h_ms, t.total_ms, t.launches, int(vis.atomic_buffer[3::4].sum())), flush=True)
