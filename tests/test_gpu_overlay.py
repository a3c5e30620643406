"""SOLVER_VIS debug overlay (gravity.c:16-111, :380-445, :550-574) through sfe_bh_overlay(): the render
buffer after the call must equal, bit for bit, what the oracle's restatement draws - and what the unmodified reference
built WITH SOLVER_VIS (oracle/_ref/libsfref_tree_vis.so) draws."""
import numpy as np
import pytest

from oracle import sfo

pytestmark = pytest.mark.gpu


def matrix(scale=0.3, a=0.7, b=0.4, tx=0.05, ty=-0.02):
    """A model-view-projection like visualization.c:363-480 builds: rotation, uniform scale, small offset (row-major)."""
    ca, sa, cb, sb = np.cos(a), np.sin(a), np.cos(b), np.sin(b)
    R = np.array([[ca, 0, sa, 0], [0, 1, 0, 0], [-sa, 0, ca, 0], [0, 0, 0, 1]]) @ \
        np.array([[1, 0, 0, 0], [0, cb, -sb, 0], [0, sb, cb, 0], [0, 0, 0, 1]])
    M = np.diag([scale, scale, scale, 1.0]) @ R
    M[0, 3] = tx; M[1, 3] = ty
    return M.astype(np.float32).ravel()


def draw_both(engine, pos, mas, box, idx, M, sx, sy, nodes_max):
    N = len(mas)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.mas[:] = mas
    engine.tree_init(sim.tree, nodes_max); sim.tree.bounds[:] = box
    assert engine.tree_build(sim.tree, N, sim.pos, sim.mas) == 0, engine.last_error()
    rng = np.random.default_rng(7)
    start = rng.random(sx * sy * 4).astype(np.float32)            # `+=` onto existing content
    rb = start.copy()
    assert engine.solve_gravity_part_tree_vis(sim, M, rb, sx, sy, idx=idx) == 0, engine.last_error()
    rc, nodes, n, *_ = sfo.tree_build_sorted(box, pos, mas, nodes_max)
    want = sfo.bh_overlay(nodes[:n], pos, mas, idx, M, start.copy(), sx, sy)
    return rb, want, start, sim


@pytest.mark.parametrize("N,sx,sy", [(20000, 640, 360), (3000, 33, 17), (50, 200, 200), (9, 64, 64)])
def test_overlay_equals_the_oracle_drawing(engine, N, sx, sy):
    pos, vel, mas = sfo.initcond(N)
    box = sfo.root_box(pos, 0)
    for idx, M in ((3, matrix()), (3, matrix(2.5, 0.1, 1.2)), (N - 1, matrix(0.05)), (min(17, N - 1), matrix(40.0, tx=-3.0))):
        rb, want, start, sim = draw_both(engine, pos, mas, box, idx, M, sx, sy, 2 * N + 64)
        assert np.array_equal(rb, want), (N, idx)
    assert not np.array_equal(rb, start) or N <= 50
    # idx outside the particle range draws nothing (gravity.c:381 never matches)
    rb2 = start.copy()
    assert engine.lib().sfe_bh_overlay(N + 5, 0.5, matrix().ctypes.data, rb2.ctypes.data, sx, sy) == 0
    assert np.array_equal(rb2, start)


def test_overlay_of_a_particle_outside_the_box_and_large_theta(engine):
    """Particle 3 far outside the root box accepts the ROOT as a monopole (its deferred moments must be folded), and with
    a huge theta nothing is opened either."""
    N = 5000
    pos, vel, mas = sfo.initcond(N)
    box = sfo.root_box(pos, 0)
    far = pos.copy(); far[9:12] = [40.0, -35.0, 50.0]
    rb, want, start, sim = draw_both(engine, far, mas, box, 3, matrix(0.01), 320, 200, 2 * N)
    assert np.array_equal(rb, want) and not np.array_equal(rb, start)
    sim2 = engine.Sim(N); sim2.pos[:] = pos; sim2.mas[:] = mas
    engine.tree_init(sim2.tree, 2 * N); sim2.tree.bounds[:] = box
    assert engine.tree_build(sim2.tree, N, sim2.pos, sim2.mas) == 0
    rb = np.zeros(320 * 200 * 4, np.float32)
    assert engine.solve_gravity_part_tree_vis(sim2, matrix(), rb, 320, 200, theta=1.0e9) == 0
    rc, nodes, n, *_ = sfo.tree_build_sorted(box, pos, mas, 2 * N)
    want = sfo.bh_overlay(nodes[:n], pos, mas, 3, matrix(), np.zeros(320 * 200 * 4, np.float32), 320, 200, theta=1.0e9)
    assert np.array_equal(rb, want)


def test_overlay_at_workload_size(engine):
    """C2 (N = 2^20) at the reference's default image size 1920 x 1080."""
    N = 1 << 20
    pos, vel, mas = sfo.initcond(N)
    box = sfo.root_box(pos, 0)
    rb, want, start, sim = draw_both(engine, pos, mas, box, 3, matrix(0.28, 0.9, 0.3), 1920, 1080, N)
    assert np.array_equal(rb, want)
    assert (rb != start).sum() > 10000


@pytest.mark.skipif(not sfo.ref_available("tree_vis"), reason="reference SOLVER_VIS library not built")
def test_overlay_equals_the_reference_built_with_solver_vis(engine):
    N, sx, sy = 30000, 800, 450
    M = matrix(0.31, 0.5, 0.8)
    r = sfo.RefSim("tree_vis", N); r.initcond()
    rb_ref = np.full(sx * sy * 4, 0.125, np.float32)
    r.solv_vis(0, M, rb_ref, sx, sy)                      # the reference's own sim_solv(sim, render_buffer, w, h)
    sim = engine.Sim(N); sim.pos[:] = r.pos; sim.mas[:] = r.mas
    engine.tree_init(sim.tree, 2 * N); sim.tree.bounds[:] = engine.root_box(sim)
    assert engine.tree_build(sim.tree, N, sim.pos, sim.mas) == 0
    rb = np.full(sx * sy * 4, 0.125, np.float32)
    assert engine.solve_gravity_part_tree_vis(sim, M, rb, sx, sy) == 0, engine.last_error()
    assert np.array_equal(sim.acc, r.acc) and np.array_equal(sim.pot, r.pot)
    assert np.array_equal(rb, rb_ref)
    r.close()


def test_overlay_errors(engine):
    L = engine.lib()
    rb = np.zeros(64, np.float32)
    M = matrix()
    assert L.sfe_bh_overlay(3, 0.5, None, rb.ctypes.data, 4, 4) == engine.STARFLOOD_FAILURE
    assert L.sfe_bh_overlay(3, 0.5, M.ctypes.data, rb.ctypes.data, 0, 4) == engine.STARFLOOD_FAILURE


def test_overlay_with_pinned_state_and_pageable_render_buffer_in_a_fresh_context(engine):
    """The first pageable copy of a context allocates the pinned bounce buffer.  When that first copy happens INSIDE
    sfe_bh_overlay() - the state block pinned with sfe_host_register (what the shims do), the render buffer pageable - the
    overlay's own buffers must survive it (an earlier version released them from ensure_bounce(): use after free)."""
    engine.shutdown(); engine.init()                              # no bounce buffer yet
    N, sx, sy = 6000, 320, 200
    pos, vel, mas = sfo.initcond(N)
    box = sfo.root_box(pos, 0)
    sim = engine.Sim(N, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas
    engine.tree_init(sim.tree, 2 * N); sim.tree.bounds[:] = box
    assert engine.tree_build(sim.tree, N, sim.pos, sim.mas) == 0, engine.last_error()
    rc, nodes, n, *_ = sfo.tree_build_sorted(box, pos, mas, 2 * N)
    start = np.random.default_rng(11).random(sx * sy * 4).astype(np.float32)
    for rep in range(3):                                          # the second and third call reuse the overlay buffers
        rb = start.copy()                                         # pageable
        assert engine.solve_gravity_part_tree_vis(sim, matrix(), rb, sx, sy) == 0, engine.last_error()
        assert np.array_equal(rb, sfo.bh_overlay(nodes[:n], pos, mas, 3, matrix(), start.copy(), sx, sy)), rep
    acc, pot, st = sfo.bh_walk(nodes[:n], pos, mas)
    assert np.array_equal(sim.acc, acc) and np.array_equal(sim.pot, pot)


def test_host_upload_of_other_particles_retires_the_tree(engine):
    """sfe_bh() walks the records sfe_tree_build() uploaded; a host PP / SPH call in between uploads other records into
    the same device buffer, so the tree is retired instead of being walked against them."""
    N = 3000
    pos, vel, mas = sfo.initcond(N)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.mas[:] = mas
    engine.tree_init(sim.tree, 2 * N); sim.tree.bounds[:] = sfo.root_box(pos, 0)
    assert engine.tree_build(sim.tree, N, sim.pos, sim.mas) == 0
    assert engine.solve_gravity_part_tree(sim) == 0
    other = engine.Sim(N); other.pos[:] = pos[::-1]; other.mas[:] = mas
    assert engine.solve_gravity_part_part(other) == 0
    assert engine.solve_gravity_part_tree(sim) == engine.STARFLOOD_FAILURE and "no tree" in engine.last_error()
    rb = np.zeros(4 * 64 * 64, np.float32)
    assert engine.lib().sfe_bh_overlay(3, 0.5, matrix().ctypes.data, rb.ctypes.data, 64, 64) == engine.STARFLOOD_FAILURE
