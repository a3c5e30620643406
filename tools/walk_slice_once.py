"""Walk of a fraction of the targets (what one rank of a multi-GPU run does): python tools/walk_slice_once.py [parts]
The particles are put in Morton order on the host first, so that an index slice is a contiguous range of the tree's slot
order - the shape of a rank's share in sfe_dev_bh_peer - and the result of every slice is hashed (SFE_WALK_SP variants
must print the same hashes)."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from starflood_b200 import engine as e, ics
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = 1 << 20
e.init()
pos, vel, mas = ics.uniform_ball(N)
pos = np.asarray(pos, dtype=np.float32).reshape(N, 3); mas = np.asarray(mas, dtype=np.float32).reshape(N)
q = np.clip(((pos - pos.min(0)) / (pos.max(0) - pos.min(0)) * 1023.0), 0, 1023).astype(np.uint64)
def spread(v):
    v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F; v = (v | (v << 4)) & 0x030C30C3; v = (v | (v << 2)) & 0x09249249
    return v
key = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
o = np.argsort(key, kind='stable'); pos = np.ascontiguousarray(pos[o]); mas = np.ascontiguousarray(mas[o])
sim = e.Sim(N, pinned=True); sim.pos[:] = pos.reshape(sim.pos.shape); sim.mas[:] = mas
e.tree_init(sim.tree, N + 64); sim.tree.bounds[:] = e.root_box(sim)
assert e.tree_build(sim.tree, N, sim.pos, sim.mas) == 0
ni = N // parts
for fl in (e.WALK_EXACT, e.WALK_FAST):
    for i0 in sorted({0, (N // parts) * (parts // 2), N - ni}):
        ts = []
        for r in range(4):
            rc = e.solve_gravity_part_tree(sim, fl, i0=i0, ni=ni); ts.append(e.timings().solve_ms)
        h = hashlib.sha1(np.ascontiguousarray(np.asarray(sim.acc).reshape(-1, 3)[i0:i0 + ni]).tobytes() + np.ascontiguousarray(sim.pot[i0:i0 + ni]).tobytes()).hexdigest()[:12]
        print('walk of N/%d targets from %7d, flavour %d: %.3f ms  SP=%s  %s' % (parts, i0, fl, min(ts[1:]), os.environ.get('SFE_WALK_SP', 'auto'), h))
