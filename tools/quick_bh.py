"""Scratch driver: tree build + walk parity and timings on the GPU box."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import sfo
from starflood_b200 import engine as e

def rel_rms(a, ref):
    a = np.asarray(a, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.sqrt(np.sum((a - ref) ** 2) / np.sum(ref ** 2)))

e.init()
sizes = [int(x) for x in sys.argv[1:]] or [100, 4096, 65536, 1048576]
for N in sizes:
    pos, vel, mas = sfo.initcond(N)
    b = sfo.root_box(pos, 0)
    maxn = N + 64
    t0 = time.time(); rc, onodes, oc, okeys, operm, onin = sfo.tree_build_sorted(b, pos, mas, maxn); t_or = time.time() - t0
    sim = e.Sim(N, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas
    e.tree_init(sim.tree, maxn); sim.tree.bounds[:] = b
    rc = e.tree_build(sim.tree, N, sim.pos, sim.mas); tb = e.timings()
    print('N', N, 'build rc', rc, 'build_ms', round(tb.build_ms, 3), 'h2d', round(tb.h2d_ms, 3), 'launches', tb.launches, 'oracle nodes', oc, 'oracle build s', round(t_or, 3))
    rc = e.tree_download(sim.tree)
    gn = sim.tree.node[:sim.tree.num_nodes]
    print('  download rc', rc, 'nodes', sim.tree.num_nodes, 'table bit-exact', gn.tobytes() == onodes[:oc].tobytes())
    if sim.tree.num_nodes == oc:
        for f in ('parent', 'bodies', 'branch', 'bounds', 'params'):
            eq = np.array_equal(gn[f], onodes[:oc][f])
            if not eq:
                bad = np.nonzero((gn[f] != onodes[:oc][f]).reshape(oc, -1).any(axis=1))[0]
                print('   ', f, 'MISMATCH at', len(bad), 'nodes, first', bad[:5], gn[f][bad[0]], onodes[f][bad[0]], 'bodies', gn['bodies'][bad[0]])
    keys, perm, nin = e.tree_keys(N)
    print('  keys', np.array_equal(keys, okeys), 'perm', np.array_equal(perm, operm), 'inbox', nin, onin)
    acc_o, pot_o, st = sfo.bh_walk(onodes[:oc], pos, mas)
    for fl, name in ((e.WALK_EXACT, 'exact'), (e.WALK_FAST, 'fast')):
        for r in range(2):
            rc = e.solve_gravity_part_tree(sim, fl); tw = e.timings()
        print('  walk', name, 'rc', rc, 'solve_ms', round(tw.solve_ms, 3), 'd2h', round(tw.d2h_ms, 3), 'bit-exact acc', np.array_equal(sim.acc, acc_o), 'pot', np.array_equal(sim.pot, pot_o),
              'relrms acc', rel_rms(sim.acc, acc_o), 'pot', rel_rms(sim.pot, pot_o))
    # repeated builds: timing
    for r in range(3):
        e.tree_build(sim.tree, N, sim.pos, sim.mas); tb = e.timings()
        e.solve_gravity_part_tree(sim, e.WALK_FAST); tw = e.timings()
        print('  rep', r, 'build_ms', round(tb.build_ms, 3), 'h2d', round(tb.h2d_ms, 3), 'walk_ms', round(tw.solve_ms, 3), 'step particle-steps/s %.3e' % (N / ((tb.build_ms + tw.solve_ms) * 1e-3)))
    # slice
    i0, ni = N // 3, N // 2
    sim.acc[:] = 0; sim.pot[:] = 0
    rc = e.solve_gravity_part_tree(sim, e.WALK_EXACT, i0=i0, ni=ni)
    print('  slice rc', rc, 'in-slice exact', np.array_equal(sim.acc[3*i0:3*(i0+ni)], acc_o[3*i0:3*(i0+ni)]), 'outside untouched', not sim.acc[:3*i0].any() and not sim.acc[3*(i0+ni):].any())
# out-of-box + tie cases
N = 5000
pos, vel, mas = sfo.initcond(N)
b = np.array([-0.5, -0.4, -0.6, 0.7, 0.8, 0.6], np.float32)   # many particles outside, non-cubic box
pos[0:3] = [0.1, 0.2, 0.0]  # exactly on mid planes: x mid = 0.1, y mid = 0.2, z mid = 0
pos[3*4000:3*4000+3] = [0.1, 0.2, 0.0]
pos[3*17:3*17+3] = [50.0, 60.0, -70.0]   # far outside: accepts the root
rc, onodes, oc, okeys, operm, onin = sfo.tree_build_sorted(b, pos, mas, 8192)
sim = e.Sim(N); sim.pos[:] = pos; sim.mas[:] = mas
e.tree_init(sim.tree, 8192); sim.tree.bounds[:] = b
print('oob build rc', e.tree_build(sim.tree, N, sim.pos, sim.mas), 'inbox', onin)
acc_o, pot_o, st = sfo.bh_walk(onodes[:oc], pos, mas)
rc = e.solve_gravity_part_tree(sim, e.WALK_EXACT)
print('oob walk exact', rc, np.array_equal(sim.acc, acc_o), np.array_equal(sim.pot, pot_o), rel_rms(sim.acc, acc_o))
e.tree_download(sim.tree)
print('oob table', sim.tree.node[:sim.tree.num_nodes].tobytes() == onodes[:oc].tobytes())
keys, perm, nin = e.tree_keys(N)
print('oob keys', np.array_equal(keys, okeys), 'perm', np.array_equal(perm, operm))
# overflow
e.tree_init(sim.tree, 64); sim.tree.bounds[:] = b
print('overflow rc (expect -1)', e.tree_build(sim.tree, N, sim.pos, sim.mas), e.last_error())
print('walk after failed build (expect -1)', e.solve_gravity_part_tree(sim, e.WALK_EXACT))
