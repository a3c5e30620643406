"""Oracle vs the golden vectors produced by the unmodified reference (tests/golden/make_golden.py).
Everything on this path is deterministic fp32 / integer work, so the bar is bit-exact."""
import os

import numpy as np
import pytest

from oracle import sfo

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def test_initcond_and_pp_bit_exact():
    g = load("pp_n1024.npz")
    pos, vel, mas = sfo.initcond(1024)
    assert np.array_equal(pos, g["pos"]) and np.array_equal(vel, g["vel"]) and np.array_equal(mas, g["mas"])
    acc, pot = sfo.pp(pos, mas)
    assert np.array_equal(acc, g["acc"])
    assert np.array_equal(pot, g["pot"])


def test_pp_f64_second_opinion():
    g = load("pp_n1024.npz")
    a64, p64 = sfo.pp_f64(g["pos"], g["mas"])
    from conftest import rel_rms
    assert rel_rms(g["acc"], a64) < 1e-6      # 5.4e-8 at N=4096
    assert rel_rms(g["pot"], p64) < 1e-6


@pytest.mark.parametrize("name,step", [("bh_n4096.npz", 0), ("bh_n4096_step7.npz", 7)])
def test_root_box_tree_and_walk_bit_exact(name, step):
    g0 = load("bh_n4096.npz")
    g = load(name)
    pos, mas = g0["pos"], g0["mas"]
    b = sfo.root_box(pos, step)
    assert np.array_equal(b, g["bounds"])
    gold = g["nodes"]
    # insertion restatement -> canonical numbering
    rc, nodes, n = sfo.tree_build_insert(b, pos, mas, len(gold) + 9)
    assert rc == 0 and n == len(gold)
    canon, nc = sfo.tree_canonicalise(nodes[:n])
    assert canon[:nc].tobytes() == gold.tobytes()
    # sort-based construction (what the CUDA engine follows) produces the canonical table directly
    rc, snodes, sn, keys, perm, nin = sfo.tree_build_sorted(b, pos, mas, len(gold) + 9)
    assert rc == 0 and sn == len(gold) and nin == 4096
    assert snodes[:sn].tobytes() == gold.tobytes()
    # keys / permutation: permutation is the stable (key, index) order
    order = np.lexsort((np.arange(4096), keys)).astype(np.uint32)
    assert np.array_equal(order, perm)
    acc, pot, st = sfo.bh_walk(gold, pos, mas)
    assert np.array_equal(acc, g["acc"]) and np.array_equal(pot, g["pot"])
    assert st.iterations == 9 * st.opened + st.evals + st.empty


def test_sph_bit_exact():
    g = load("sph_n1024.npz")
    pos, mas, rad = g["pos"], g["mas"], g["rad"]
    acc, pot = sfo.pp(pos, mas)
    rho = sfo.sph_density(pos, mas, rad)
    prs = sfo.sph_eos(rho)
    acc = sfo.sph_force(pos, mas, rad, rho, prs, acc)
    assert np.array_equal(rho, g["rho"])
    assert np.array_equal(prs, g["prs"])
    assert np.array_equal(acc, g["acc"])
    assert np.array_equal(pot, g["pot"])


def test_kdk_energy_series_bit_exact():
    g = load("energy_pp_n512.npz")
    N = 512
    pos, vel, mas = sfo.initcond(N)
    acc, pot = sfo.pp(pos, mas)
    L = sfo.lib()
    for s in range(40):
        L.sfo_kick(N, vel, acc, 1.0); L.sfo_drift(N, pos, vel, 1.0)
        acc, pot = sfo.pp(pos, mas)
        L.sfo_kick(N, vel, acc, 1.0)
        T, V = sfo.energies(vel, mas, pot)
        assert abs(T - g["T"][s]) <= 1e-12 * abs(g["T"][s]) and abs(V - g["V"][s]) <= 1e-12 * abs(g["V"][s])
    assert np.array_equal(pos, g["pos"]) and np.array_equal(vel, g["vel"])


def test_tree_edge_cases():
    # empty, single particle, exactly 8 and 9 bodies, all particles outside the box, coincident particles
    b = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    for n in (0, 1, 8, 9):
        pos = (np.linspace(-0.9, 0.9, 3 * max(n, 1))[: 3 * n]).astype(np.float32)
        mas = np.full(n, 0.5, np.float32)
        rc, nodes, cnt = sfo.tree_build_insert(b, pos, mas, 64)
        rc2, snodes, scnt, keys, perm, nin = sfo.tree_build_sorted(b, pos, mas, 64)
        assert rc == 0 and rc2 == 0 and cnt == scnt == (9 if n > 8 else 1) and nin == n
        canon, nc = sfo.tree_canonicalise(nodes[:cnt])
        assert canon[:nc].tobytes() == snodes[:scnt].tobytes()
    pos = np.full(30, 5.0, np.float32); mas = np.ones(10, np.float32)
    rc, snodes, scnt, keys, perm, nin = sfo.tree_build_sorted(b, pos, mas, 64)
    assert rc == 0 and scnt == 1 and nin == 0 and (keys == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    # nine coincident particles can never be separated: depth overflow like tree.c:186-189
    pos = np.full(27, 0.3, np.float32); mas = np.ones(9, np.float32)
    rc, nodes, cnt = sfo.tree_build_insert(b, pos, mas, 4096)
    rc2 = sfo.tree_build_sorted(b, pos, mas, 4096)[0]
    assert rc == -1 and rc2 == -1
    # node budget: success iff the final count stays below max_nodes (tree.c:180-183)
    pos, vel, mas = sfo.initcond(512)
    bb = sfo.root_box(pos, 0)
    rc, nodes, cnt = sfo.tree_build_insert(bb, pos, mas, 4096)
    assert rc == 0
    assert sfo.tree_build_insert(bb, pos, mas, cnt + 1)[0] == 0 and sfo.tree_build_sorted(bb, pos, mas, cnt + 1)[0] == 0
    assert sfo.tree_build_insert(bb, pos, mas, cnt)[0] == -1 and sfo.tree_build_sorted(bb, pos, mas, cnt)[0] == -1


def test_tie_rule_matters():
    # a particle exactly on a mid-plane goes low if it is among the first 8 members, high otherwise
    b = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    rng = np.random.default_rng(1)
    pos = rng.uniform(-0.9, 0.9, 3 * 64).astype(np.float32)
    pos[0:3] = 0.0          # resident of the root when it splits -> octant 0
    pos[3 * 40:3 * 40 + 3] = 0.0   # arrives later -> octant 7
    mas = np.ones(64, np.float32)
    rc, snodes, scnt, keys, perm, nin = sfo.tree_build_sorted(b, pos, mas, 512)
    assert (int(keys[0]) >> 45) == 0 and (int(keys[40]) >> 45) == 7
    rc, nodes, cnt = sfo.tree_build_insert(b, pos, mas, 512)
    canon, nc = sfo.tree_canonicalise(nodes[:cnt])
    assert canon[:nc].tobytes() == snodes[:scnt].tobytes()


def test_vis_draw_bit_exact():
    """vis_draw() (visualization.c:253-739): matMVP, the integer accumulation buffer and the render buffer of the
    reference's own run (tests/golden/vis_n4096.npz; the image is narrower than the projected ball, so samples are
    dropped at the border)."""
    g = load("vis_n4096.npz")
    sx, sy, step = int(g["sizex"]), int(g["sizey"]), int(g["step"])
    rbar, mvp, ab, rb = sfo.vis_draw(np.ascontiguousarray(g["pos"]), np.ascontiguousarray(g["pot"]), step, sx, sy)
    assert np.array_equal(mvp, g["mvp"])
    assert np.array_equal(ab, g["atomic"])
    assert np.array_equal(rb, g["render"])
    assert 0 < int(ab[3::4].sum()) < 128 * len(g["pot"])
