"""BASELINE.json's larger configurations as parity cases: C4 = PP N = 2^22 sharded by target slice, C5 =
Barnes-Hut N = 2^24 galaxy-pair collision, tree per GPU.  One GPU evaluates what one rank of an 8-GPU box would; the
checks are the oracle on sampled targets plus size-independent properties."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_rms
from oracle import sfo

pytestmark = pytest.mark.gpu


def test_c4_pp_one_rank_of_eight(engine):
    """N = 2^22: rank 3 of 8 evaluates its target slice over all sources (5.5e11 interactions)."""
    N, world, rank = 1 << 22, 8, 3
    pos, vel, mas = sfo.initcond(N)
    i0, i1 = engine.sim_slice(N, rank, world)
    sim = engine.Sim(N, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas
    sim.acc[:] = 7.0; sim.pot[:] = 7.0
    assert engine.lib().sfe_pp_slice(sim.pos.ctypes.data, sim.mas.ctypes.data, sim.acc.ctypes.data, sim.pot.ctypes.data,
                                     N, i0, i1 - i0, engine.G, engine.EPS2) == 0, engine.last_error()
    # rows of other ranks are not touched
    assert np.all(sim.acc[:3 * i0] == 7.0) and np.all(sim.acc[3 * i1:] == 7.0) and np.all(sim.pot[:i0] == 7.0)
    for a in (i0, (i0 + i1) // 2, i1 - 96):
        acc, pot = sfo.pp(pos, mas, a, a + 96)
        assert rel_rms(sim.acc[3 * a:3 * a + 288], acc[3 * a:3 * a + 288]) <= 1e-5
        assert rel_rms(sim.pot[a:a + 96], pot[a:a + 96]) <= 1e-5
    own = sim.acc[3 * i0:3 * i1]
    assert np.isfinite(own).all() and (sim.pot[i0:i1] < 0).all()
    # a uniform ball pulls every particle inwards: acc . pos < 0 for (almost) all of them
    dot = np.einsum("ij,ij->i", own.reshape(-1, 3).astype(np.float64), pos[3 * i0:3 * i1].reshape(-1, 3).astype(np.float64))
    assert (dot < 0).mean() > 0.999


def test_c5_bh_galaxy_pair_full_size(engine):
    """N = 2^24 galaxy pair (the author's disabled sketch, initcond.c:265-329): the device tree equals
    the oracle's table bit for bit, particle counts are conserved, depth stays under TREE_DEPTH_MAX, and one rank's
    walk of its slice is bit-identical to the oracle walk on sampled targets."""
    N, world, rank = 1 << 24, 8, 5
    nodes_max = 1 << 23
    pos, vel, mas = sfo.initcond_galaxy_pair(N)
    b = sfo.root_box(pos, 0)
    sim = engine.Sim(N, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas
    engine.tree_init(sim.tree, nodes_max); sim.tree.bounds[:] = b
    assert engine.tree_build(sim.tree, N, sim.pos, sim.mas) == 0, engine.last_error()
    i0, i1 = engine.sim_slice(N, rank, world)
    assert engine.solve_gravity_part_tree(sim, engine.WALK_EXACT, i0=i0, ni=i1 - i0) == 0, engine.last_error()
    assert engine.tree_download(sim.tree) == 0, engine.last_error()
    n = sim.tree.num_nodes
    nodes = sim.tree.node[:n]
    # structure: the root holds every in-box particle, every internal node's children add up to it
    inbox = int(np.sum((pos.reshape(-1, 3) >= b[:3]).all(axis=1) & (pos.reshape(-1, 3) <= b[3:]).all(axis=1)))
    assert nodes["bodies"][0] == inbox
    internal = np.nonzero(nodes["bodies"] > 8)[0]
    kids = nodes["branch"][internal]
    assert np.array_equal(nodes["bodies"][kids].sum(axis=1, dtype=np.uint64), nodes["bodies"][internal].astype(np.uint64))
    assert np.array_equal(nodes["parent"][kids], np.repeat(internal[:, None], 8, axis=1).astype(np.uint32))
    # the oracle's construction of the same tree (restated from tree.c:74-318) is the same table
    rc, onodes, oc, keys, perm, nin = sfo.tree_build_sorted(b, pos, mas, nodes_max)
    assert rc == 0 and oc == n and nin == inbox
    assert onodes[:oc].tobytes() == nodes.tobytes()
    # sampled targets of the rank's slice against the oracle walk
    for a in (i0, i0 + 1234567, i1 - 2048):
        acc, pot, st = sfo.bh_walk(onodes[:oc], pos, mas, a, a + 2048)
        assert np.array_equal(sim.acc[3 * a:3 * (a + 2048)], acc[3 * a:3 * (a + 2048)])
        assert np.array_equal(sim.pot[a:a + 2048], pot[a:a + 2048])
    assert np.isfinite(sim.acc[3 * i0:3 * i1]).all()
