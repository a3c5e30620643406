"""Barnes-Hut walk parity (solve_gravity_part_tree, gravity.c:245-663) through the C ABI.  Gate from
BASELINE.json: acc within 1e-4 rel-RMS of the reference's BH at the same theta; the EXACT flavour is held to
the stronger bar it was built for: bit-identical acc and pot."""
import os

import numpy as np
import pytest

from conftest import rel_rms
from oracle import sfo

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(params=[0, 1, 2], ids=["walk_kernel", "walk_sp_kernel", "walk_p_kernel"])
def both_kernels(engine, request):
    """The walk has a throughput kernel, a latency (software-pipelined) kernel and a throughput kernel on depth-first node
    records, chosen by the number of targets and the flavour; the tests that use this fixture run with each of them forced,
    whatever the size."""
    engine.walk_select(request.param)
    yield request.param
    engine.walk_select(-1)


def solve(engine, pos, mas, bounds, flavour, theta=0.5, max_nodes=None, **kw):
    N = len(mas)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.mas[:] = mas
    engine.tree_init(sim.tree, max_nodes or N + 64); sim.tree.bounds[:] = bounds
    assert engine.tree_build(sim.tree, N, sim.pos, sim.mas) == 0, engine.last_error()
    assert engine.solve_gravity_part_tree(sim, flavour, theta=theta, **kw) == 0, engine.last_error()
    return sim


@pytest.mark.parametrize("name", ["bh_n4096.npz", "bh_n4096_step7.npz"])
def test_golden(engine, both_kernels, name):
    g0 = np.load(os.path.join(G, "bh_n4096.npz")); g = np.load(os.path.join(G, name))
    sim = solve(engine, g0["pos"], g0["mas"], g["bounds"], engine.WALK_EXACT)
    assert np.array_equal(sim.acc, g["acc"]) and np.array_equal(sim.pot, g["pot"])
    sim = solve(engine, g0["pos"], g0["mas"], g["bounds"], engine.WALK_FAST)
    assert rel_rms(sim.acc, g["acc"]) <= 1e-4 and rel_rms(sim.pot, g["pot"]) <= 1e-4


@pytest.mark.parametrize("N", [1, 9, 100, 5000, 65536])
@pytest.mark.parametrize("theta", [0.5, 0.2, 1.0, 2.5])
def test_sizes_and_thetas(engine, both_kernels, N, theta):
    if N == 65536 and theta == 0.2:
        pytest.skip("slow on the CPU oracle")
    pos, vel, mas = sfo.initcond(N)
    b = sfo.root_box(pos, 0)
    rc, onodes, oc, *_ = sfo.tree_build_sorted(b, pos, mas, N + 64)
    acc, pot, st = sfo.bh_walk(onodes[:oc], pos, mas, theta=np.float32(theta))
    sim = solve(engine, pos, mas, b, engine.WALK_EXACT, theta=theta)
    assert np.array_equal(sim.acc, acc) and np.array_equal(sim.pot, pot)      # theta 2.5 accepts the root itself
    sim = solve(engine, pos, mas, b, engine.WALK_FAST, theta=theta)
    assert rel_rms(sim.pot, pot) <= 1e-4
    if N > 1:
        assert rel_rms(sim.acc, acc) <= 1e-4


def test_out_of_box_targets_and_self_subtraction(engine, both_kernels):
    N = 5000
    pos, vel, mas = sfo.initcond(N)
    b = np.array([-0.5, -0.4, -0.6, 0.7, 0.8, 0.6], np.float32)        # elongated box, 58 % of particles outside
    pos[3 * 17:3 * 17 + 3] = [50.0, 60.0, -70.0]                       # accepts the root as a monopole
    rc, onodes, oc, *_ = sfo.tree_build_sorted(b, pos, mas, 8192)
    acc, pot, st = sfo.bh_walk(onodes[:oc], pos, mas)
    assert st.self_sub > 0
    sim = solve(engine, pos, mas, b, engine.WALK_EXACT, max_nodes=8192)
    assert np.array_equal(sim.acc, acc) and np.array_equal(sim.pot, pot)


@pytest.mark.parametrize("theta", [0.5, 2.0])
def test_targets_on_cell_faces_and_large_theta(engine, both_kernels, theta):
    """gravity.c:473-476 tests `lo <= r <= hi` on BOTH sides, so a target that lies exactly on a face shared by two cells is
    taken out of the monopole of a leaf it does not belong to; with theta = 2 accepted INTERNAL cells contain their target
    too.  The walk's filters in front of the bounds test (a cube about a leaf's COM, a sphere about a cell's centre) must let
    all of these through: positions on a dyadic lattice inside a dyadic root box put thousands of targets on cell faces."""
    rng = np.random.default_rng(7)
    N = 6000
    idx = rng.choice(32 ** 3, N, replace=False)
    g = np.stack([idx % 32, (idx // 32) % 32, idx // 1024], 1).astype(np.float32)
    pos = (g / 16.0 - 1.0).astype(np.float32)                          # multiples of 1/16 in [-1, 1)
    pos[::7] += rng.normal(0, 0.01, (len(pos[::7]), 3)).astype(np.float32)
    pos = np.ascontiguousarray(pos.reshape(-1))
    mas = ((rng.random(N).astype(np.float32) + 0.5) / N).astype(np.float32)
    b = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    rc, onodes, oc, *_ = sfo.tree_build_sorted(b, pos, mas, 4 * N)
    acc, pot, st = sfo.bh_walk(onodes[:oc], pos, mas, theta=theta)
    assert st.self_sub > 1.5 * N                                       # more subtractions than targets: the shared faces
    sim = solve(engine, pos, mas, b, engine.WALK_EXACT, theta=theta, max_nodes=4 * N)
    assert np.array_equal(sim.acc, acc) and np.array_equal(sim.pot, pot)
    sim = solve(engine, pos, mas, b, engine.WALK_FAST, theta=theta, max_nodes=4 * N)
    assert rel_rms(sim.acc, acc) <= 1e-4 and rel_rms(sim.pot, pot) <= 1e-4


def test_root_box_that_is_a_point(engine, both_kernels):
    """A root box of zero extent (three coincident particles inside it, everything else outside): a one-leaf tree whose depth
    cannot be read off the extents - the depth-first numbering raises its flag and walk_p_kernel runs walk_kernel's loop."""
    N = 300
    pos, vel, mas = sfo.initcond(N)
    pos[3 * 7:3 * 7 + 3] = pos[0:3]; pos[3 * 9:3 * 9 + 3] = pos[0:3]
    b = np.array([pos[0], pos[1], pos[2], pos[0], pos[1], pos[2]], np.float32)
    rc, onodes, oc, *_ = sfo.tree_build_sorted(b, pos, mas, 1024)
    acc, pot, st = sfo.bh_walk(onodes[:oc], pos, mas)
    assert oc == 1 and st.self_sub == 3
    sim = solve(engine, pos, mas, b, engine.WALK_EXACT, max_nodes=1024)
    assert np.array_equal(sim.acc, acc) and np.array_equal(sim.pot, pot)


def test_body_at_the_very_centre_of_the_root_box(engine, both_kernels):
    """d^2 = 0 makes the reference's opening ratio 0 (gravity.c:377): the one in-box point that does NOT open the root but takes it
    as a monopole (minus itself).  At this size the build's top-level pass tells the host whether the root's deferred fold can be
    needed at all (every body in the box, theta * rin2 / w^2 < 1): a body at the exact centre must keep the gate's pre-pass."""
    N = 70000                                                          # (>= 65536: the one-pass top levels)
    pos, vel, mas = sfo.initcond(N)
    b = np.array([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5], np.float32)        # every body inside, centre (0, 0, 0)
    pos[3 * 1234:3 * 1234 + 3] = [0.0, 0.0, 0.0]
    rc, onodes, oc, _, _, nin = sfo.tree_build_sorted(b, pos, mas, 2 * N)
    assert nin == N
    acc, pot, st = sfo.bh_walk(onodes[:oc], pos, mas)
    sim = solve(engine, pos, mas, b, engine.WALK_EXACT, max_nodes=2 * N)
    assert np.array_equal(sim.acc, acc) and np.array_equal(sim.pot, pot)
    # ... and without it the gate is closed on the host: same parity, no root fold
    pos[3 * 1234:3 * 1234 + 3] = [0.01, 0.0, 0.0]
    rc, onodes, oc, *_ = sfo.tree_build_sorted(b, pos, mas, 2 * N)
    acc, pot, st = sfo.bh_walk(onodes[:oc], pos, mas)
    sim = solve(engine, pos, mas, b, engine.WALK_EXACT, max_nodes=2 * N)
    assert np.array_equal(sim.acc, acc) and np.array_equal(sim.pot, pot)


def test_distances_beyond_the_newton_range(engine, both_kernels):
    """q = r^2 + eps^2 >= 2^16 leaves the range in which the EXACT walk's Newton sequences were verified: those sources go
    through the IEEE forms (walk_kernel: per source; walk_sp_kernel: the loop without RANGE_OK for warps that can meet one)."""
    N = 3000
    pos, vel, mas = sfo.initcond(N)
    b = sfo.root_box(pos, 0)
    pos[3 * 5:3 * 5 + 3] = [300.0, -200.0, 150.0]                      # r^2 = 152 500 to everything else
    pos[3 * 2100:3 * 2100 + 3] = [-180.0, 10.0, 0.5]                   # r^2 = 32 500: inside the range, outside the warp test
    rc, onodes, oc, *_ = sfo.tree_build_sorted(b, pos, mas, N + 64)
    acc, pot, st = sfo.bh_walk(onodes[:oc], pos, mas)
    sim = solve(engine, pos, mas, b, engine.WALK_EXACT)
    assert np.array_equal(sim.acc, acc) and np.array_equal(sim.pot, pot)
    # ... and with the sources far away too: a root box around everything
    lo = pos.reshape(-1, 3).min(0) - 1; hi = pos.reshape(-1, 3).max(0) + 1
    b2 = np.concatenate([lo, hi]).astype(np.float32)
    rc, onodes, oc, *_ = sfo.tree_build_sorted(b2, pos, mas, N + 64)
    acc, pot, st = sfo.bh_walk(onodes[:oc], pos, mas)
    sim = solve(engine, pos, mas, b2, engine.WALK_EXACT)
    assert np.array_equal(sim.acc, acc) and np.array_equal(sim.pot, pot)


def test_slices_tile_the_answer(engine, both_kernels):
    N = 30000
    pos, vel, mas = sfo.initcond(N)
    b = sfo.root_box(pos, 0)
    full = solve(engine, pos, mas, b, engine.WALK_EXACT)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.mas[:] = mas
    engine.tree_init(sim.tree, N + 64); sim.tree.bounds[:] = b
    assert engine.tree_build(sim.tree, N, sim.pos, sim.mas) == 0
    for i0, ni in ((0, 10000), (10000, 7), (10007, 19993)):
        assert engine.solve_gravity_part_tree(sim, engine.WALK_EXACT, i0=i0, ni=ni) == 0
    assert np.array_equal(sim.acc, full.acc) and np.array_equal(sim.pot, full.pot)


def test_needs_matching_tree(engine):
    sim = engine.Sim(64); sim.mas[:] = 1
    engine.tree_init(sim.tree, 64); sim.tree.bounds[:] = [-1, -1, -1, 1, 1, 1]
    assert engine.tree_build(sim.tree, 64, sim.pos, sim.mas) in (0, -1)
    other = engine.Sim(65); other.mas[:] = 1
    assert engine.solve_gravity_part_tree(other) == engine.STARFLOOD_FAILURE


def test_full_size(engine):
    """C2: N = 2^20, theta = 0.5, root box of step 0: every particle bit-identical to the oracle walk
    (which reproduces the reference bitwise)."""
    N = 1 << 20
    pos, vel, mas = sfo.initcond(N)
    b = sfo.root_box(pos, 0)
    rc, onodes, oc, *_ = sfo.tree_build_sorted(b, pos, mas, N)
    acc, pot, st = sfo.bh_walk(onodes[:oc], pos, mas)
    sim = solve(engine, pos, mas, b, engine.WALK_EXACT, max_nodes=N)
    assert np.array_equal(sim.acc, acc) and np.array_equal(sim.pot, pot)
    fast = solve(engine, pos, mas, b, engine.WALK_FAST, max_nodes=N)
    assert rel_rms(fast.acc, acc) <= 1e-4 and rel_rms(fast.pot, pot) <= 1e-4
    try:                                       # the latency kernel on the same tree: the same bits, both flavours
        engine.walk_select(1)
        sp = solve(engine, pos, mas, b, engine.WALK_EXACT, max_nodes=N)
        assert np.array_equal(sp.acc, acc) and np.array_equal(sp.pot, pot)
        spf = solve(engine, pos, mas, b, engine.WALK_FAST, max_nodes=N)
        assert np.array_equal(spf.acc, fast.acc) and np.array_equal(spf.pot, fast.pot)
    finally:
        engine.walk_select(-1)
    # walk statistics that the roofline model uses: V = opened + evals + empty = 514.3
    assert abs((st.opened + st.evals + st.empty) / N - 514.3) < 0.1


def test_exact_math_exhaustive(engine):
    """The EXACT flavour's 1/q, sqrt(q), 1/sqrt(q) sequences give the IEEE result for every float in [2^-24, 2^16)."""
    assert engine.selftest_exact_math() == [0, 0, 0]


@pytest.mark.parametrize("N,step", [(9, 0), (20000, 0), (20000, 5), (200000, 2)])
def test_solv_tree_is_the_tree_branch_of_sim_solv(engine, N, step):
    """sfe_solv_tree(): root box of step_number + tree_build + walk in one call on host arrays (simulation.c:245-325) -
    the box equals the reference's bit for bit, acc / pot equal the oracle's walk on the oracle's tree."""
    pos, vel, mas = sfo.initcond(N)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.mas[:] = mas; sim.step_number = step
    engine.tree_init(sim.tree, N + 64)
    assert engine.solv_tree(sim) == 0, engine.last_error()
    b = sfo.root_box(pos, step)
    assert np.array_equal(sim.tree.bounds, b)
    rc, onodes, oc, *_ = sfo.tree_build_sorted(b, pos, mas, N + 64)
    acc, pot, st = sfo.bh_walk(onodes[:oc], pos, mas)
    assert np.array_equal(sim.acc, acc) and np.array_equal(sim.pot, pot)


def test_solv_tree_reports_a_failed_build_like_the_reference(engine):
    N = 20000
    pos, vel, mas = sfo.initcond(N)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.mas[:] = mas
    engine.tree_init(sim.tree, 1000)                                   # far too few nodes (tree.c:180-183)
    assert engine.solv_tree(sim) == engine.STARFLOOD_FAILURE
    assert "Exceeded maximum number of tree nodes" in engine.last_error()
    engine.tree_init(sim.tree, N + 64)
    assert engine.solv_tree(sim) == 0, engine.last_error()             # and the engine is usable afterwards
