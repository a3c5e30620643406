"""Octree build parity (tree_build, tree.c:74-318) through the C ABI: node table (cell set, bodies, leaf member
lists, bounds, moments) bit-exact against the reference-validated oracle and the golden table dumped from the
reference; Morton keys and sort permutation bit-exact against the oracle restatement."""
import os

import numpy as np
import pytest

from oracle import sfo

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def build(engine, pos, mas, bounds, max_nodes, depth_max=16):
    N = len(mas)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.mas[:] = mas
    engine.tree_init(sim.tree, max_nodes); sim.tree.bounds[:] = bounds
    rc = engine.tree_build(sim.tree, N, sim.pos, sim.mas, depth_max)
    return rc, sim


def check_against_oracle(engine, pos, mas, bounds, max_nodes, depth_max=16):
    rc, onodes, oc, okeys, operm, onin = sfo.tree_build_sorted(bounds, pos, mas, max_nodes, depth_max)
    grc, sim = build(engine, pos, mas, bounds, max_nodes, depth_max)
    assert grc == rc
    if rc != 0:
        return None
    assert engine.tree_download(sim.tree) == 0
    assert sim.tree.num_nodes == oc
    got = sim.tree.node[:oc]
    for f in ("parent", "bodies", "branch", "bounds", "params"):
        assert got[f].tobytes() == onodes[:oc][f].tobytes(), f
    keys, perm, nin = engine.tree_keys(len(mas))
    assert nin == onin and np.array_equal(keys, okeys) and np.array_equal(perm, operm)
    return sim


def test_golden_table(engine):
    g = np.load(os.path.join(G, "bh_n4096.npz"))
    for name in ("bh_n4096.npz", "bh_n4096_step7.npz"):
        gg = np.load(os.path.join(G, name))
        rc, sim = build(engine, g["pos"], g["mas"], gg["bounds"], len(gg["nodes"]) + 9)
        assert rc == 0 and engine.tree_download(sim.tree) == 0
        assert sim.tree.node[:sim.tree.num_nodes].tobytes() == gg["nodes"].tobytes()


@pytest.mark.parametrize("N", [0, 1, 8, 9, 100, 5000, 65536])
def test_sizes(engine, N):
    pos, vel, mas = sfo.initcond(N) if N else (np.zeros(0, np.float32),) * 3
    b = sfo.root_box(pos, 0) if N else np.array([-1, -1, -1, 1, 1, 1], np.float32)
    check_against_oracle(engine, pos, mas, b, N + 64)


def test_out_of_box_ties_and_non_cubic_box(engine):
    N = 5000
    pos, vel, mas = sfo.initcond(N)
    b = np.array([-0.5, -0.375, -0.625, 0.75, 0.875, 0.625], np.float32)   # mid planes 0.125, 0.25, 0 are exact in fp32
    pos[0:3] = [0.125, 0.25, 0.0]              # exactly on the root's mid planes, resident -> low side
    pos[3 * 4000:3 * 4000 + 3] = [0.125, 0.25, 0.0]  # same point arriving later -> high side
    pos[3 * 17:3 * 17 + 3] = [50.0, 60.0, -70.0]
    mas = (mas * np.linspace(0.5, 1.5, N)).astype(np.float32)
    sim = check_against_oracle(engine, pos, mas, b, 8192)
    keys, perm, nin = engine.tree_keys(N)
    assert nin < N and keys[17] == np.uint64(0xFFFFFFFFFFFFFFFF)
    assert (int(keys[0]) >> 45) == 0 and (int(keys[4000]) >> 45) == 7


def test_all_outside(engine):
    pos = np.full(300, 9.0, np.float32); mas = np.ones(100, np.float32)
    check_against_oracle(engine, pos, mas, np.array([-1, -1, -1, 1, 1, 1], np.float32), 64)


def test_clustered_deep_tree(engine):
    rng = np.random.default_rng(3)
    N = 20000
    pos = (rng.normal(size=(N, 3)) * 1e-3).astype(np.float32)
    pos[: N // 2] += 0.7
    mas = rng.uniform(0.1, 1.0, N).astype(np.float32)
    check_against_oracle(engine, pos.ravel(), mas, np.array([-1, -1, -1, 1, 1, 1], np.float32), 2 * N)


def test_failures_like_reference(engine):
    pos, vel, mas = sfo.initcond(4096)
    b = sfo.root_box(pos, 0)
    rc, onodes, oc, *_ = sfo.tree_build_sorted(b, pos, mas, 8192)
    assert build(engine, pos, mas, b, oc + 1)[0] == 0                      # final count must stay below max_nodes
    assert build(engine, pos, mas, b, oc)[0] == engine.STARFLOOD_FAILURE  # tree.c:180-183
    assert "Exceeded maximum number of tree nodes" in engine.last_error()
    pos9 = np.full(27, 0.3, np.float32); m9 = np.ones(9, np.float32)
    assert build(engine, pos9, m9, np.array([-1, -1, -1, 1, 1, 1], np.float32), 4096)[0] == engine.STARFLOOD_FAILURE
    assert "Exceeded maximum tree depth" in engine.last_error()            # tree.c:186-189
    rc, sim = build(engine, pos9, m9, np.array([-1, -1, -1, 1, 1, 1], np.float32), 4096)
    assert engine.solve_gravity_part_tree(sim) == engine.STARFLOOD_FAILURE   # no walk on a failed build
    # shallow depth limit
    check_against_oracle(engine, pos, mas, b, 8192, depth_max=3)


def test_full_size(engine):
    """N = 2^20 (C2): bit-exact table / keys / permutation + structural properties."""
    N = 1 << 20
    pos, vel, mas = sfo.initcond(N)
    b = sfo.root_box(pos, 0)
    sim = check_against_oracle(engine, pos, mas, b, N)
    assert sim.tree.num_nodes == 400761
    nodes = sim.tree.node[:sim.tree.num_nodes]
    keys, perm, nin = engine.tree_keys(N)
    assert np.array_equal(np.sort(perm), np.arange(N, dtype=np.uint32))    # a permutation
    assert (np.diff(keys[perm].astype(np.uint64)) >= 0).all()               # sorted by key
    leaves = nodes[nodes["bodies"] <= 8]
    assert int(leaves["bodies"].sum()) == nin == int(nodes["bodies"][0])
    assert len(nodes) == 1 + 8 * int((nodes["bodies"] > 8).sum())


@pytest.mark.parametrize("sub_max,xl_min,top_min", [(8, None, None), (40, None, None), (300, 64, None), (2048, 64, None),
                                                    (8, None, 0), (40, 64, 0), (2048, None, 0)])
def test_build_variants_in_a_fresh_process(sub_max, xl_min, top_min):
    """The build has three parts: the upper four levels in one pass over the particles (sfe_top.cuh; builds of at least
    SFE_TOP_MIN particles, 65536 by default), grid-wide passes level by level below them, and one CTA per subtree root at
    the bottom (SFE_SUBTREE_MAX bodies at most; 8 switches the subtree path off).  The library reads the thresholds once,
    so every setting runs this file's parity cases in a process of its own: all passes grid-wide (8), subtrees of a few
    dozen bodies so that small trees mix the parts over several levels (40, 300), the exact chunked fold of very long
    nodes lowered to 64 members (SFE_XL_MIN) so that small trees exercise it next to the subtrees, and the one-pass top
    levels for every size (SFE_TOP_MIN=0: trees that end inside the top levels, out-of-box particles, the fall-back to
    the level-by-level passes on an exact tie)."""
    if os.environ.get("SFE_TREE_VARIANT_CHILD"):
        pytest.skip("already inside a variant run")
    import subprocess, sys
    env = dict(os.environ, SFE_SUBTREE_MAX=str(sub_max), SFE_TREE_VARIANT_CHILD="1")
    if xl_min:
        env["SFE_XL_MIN"] = str(xl_min)
    env["SFE_TOP_MIN"] = str(top_min) if top_min is not None else str(1 << 30)
    sel = "golden or sizes or ties or outside or clustered or failures" + (" or full_size" if sub_max in (8, 300) else "")
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k", sel],
                       capture_output=True, text=True, env=env, timeout=1200)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
