"""Oracle vs the unmodified reference libraries (oracle/_ref/, built by oracle/build_ref.sh from
/root/reference/src).  Skipped when the libraries are absent."""
import numpy as np
import pytest

from oracle import sfo

needs_ref = pytest.mark.skipif(not (sfo.ref_available("pp") and sfo.ref_available("tree") and sfo.ref_available("pp_sph")),
                               reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("N", [257, 4096])
def test_pp_matches_reference(N):
    r = sfo.RefSim("pp", N); r.initcond(); r.solv(0)
    pos, vel, mas = sfo.initcond(N)
    assert np.array_equal(pos, r.pos) and np.array_equal(vel, r.vel) and np.array_equal(mas, r.mas)
    acc, pot = sfo.pp(pos, mas)
    assert np.array_equal(acc, r.acc) and np.array_equal(pot, r.pot)


@needs_ref
@pytest.mark.parametrize("N,step", [(1000, 0), (65536, 3)])
def test_tree_and_walk_match_reference(N, step):
    r = sfo.RefSim("tree", N); r.initcond(); r.solv(step)
    pos, vel, mas = sfo.initcond(N)
    b = sfo.root_box(pos, step)
    ref_nodes = r.nodes().copy()
    assert np.array_equal(b, np.array(ref_nodes[0]["bounds"]))
    rc, nodes, n = sfo.tree_build_insert(b, pos, mas, len(ref_nodes) + 9)
    assert rc == 0 and nodes[:n].tobytes() == ref_nodes.tobytes()           # same numbering, same bits
    rc, snodes, sn, keys, perm, nin = sfo.tree_build_sorted(b, pos, mas, len(ref_nodes) + 9)
    canon, nc = sfo.tree_canonicalise(ref_nodes)
    assert rc == 0 and snodes[:sn].tobytes() == canon[:nc].tobytes()
    acc, pot, st = sfo.bh_walk(snodes[:sn], pos, mas)
    assert np.array_equal(acc, r.acc) and np.array_equal(pot, r.pot)
    r.close()


@needs_ref
def test_sph_matches_reference():
    N = 1500
    r = sfo.RefSim("pp_sph", N); r.initcond(); r.solv(0)
    pos, vel, mas, rad = sfo.initcond(N, sph=True)
    acc, pot = sfo.pp(pos, mas)
    rho = sfo.sph_density(pos, mas, rad); prs = sfo.sph_eos(rho)
    acc = sfo.sph_force(pos, mas, rad, rho, prs, acc)
    assert np.array_equal(rho, r.rho) and np.array_equal(prs, r.prs) and np.array_equal(acc, r.acc)


@needs_ref
def test_walk_statistics_match_survey():
    # SURVEY App. C, N = 65536: 795.2 iterations, 274.2 evaluations, 49.6 opened, 74.2 empty visits per particle
    N = 65536
    pos, vel, mas = sfo.initcond(N)
    b = sfo.root_box(pos, 0)
    rc, nodes, n, keys, perm, nin = sfo.tree_build_sorted(b, pos, mas, N)
    assert n == 42545
    acc, pot, st = sfo.bh_walk(nodes[:n], pos, mas)
    assert abs(st.iterations / N - 795.2) < 0.1 and abs(st.evals / N - 274.2) < 0.1
    assert abs(st.opened / N - 49.6) < 0.1 and abs(st.empty / N - 74.2) < 0.1


@pytest.mark.skipif(not sfo.ref_available("tree_vis"), reason="reference SOLVER_VIS library not built")
def test_overlay_restatement_equals_reference_drawing():
    """sfo_bh_overlay() against the unmodified reference compiled WITH SOLVER_VIS 3 (config.h:32): the
    render buffer after sim_solv(sim, render_buffer, sizex, sizey) is reproduced bit for bit."""
    N, sx, sy = 20000, 640, 360
    a, b, sc = 0.7, 0.4, 0.3
    ca, sa, cb, sb = np.cos(a), np.sin(a), np.cos(b), np.sin(b)
    R = np.array([[ca, 0, sa, 0], [0, 1, 0, 0], [-sa, 0, ca, 0], [0, 0, 0, 1]]) @ \
        np.array([[1, 0, 0, 0], [0, cb, -sb, 0], [0, sb, cb, 0], [0, 0, 0, 1]])
    M = (np.diag([sc, sc, sc, 1.0]) @ R).astype(np.float32).ravel()
    M[3] = 0.05; M[7] = -0.02
    r = sfo.RefSim("tree_vis", N); r.initcond()
    for step in (0, 5):
        rb = np.full(sx * sy * 4, 0.25, np.float32)
        r.solv_vis(step, M, rb, sx, sy)
        box = sfo.root_box(r.pos.copy(), step)
        rc, nodes, n, *_ = sfo.tree_build_sorted(box, r.pos.copy(), r.mas.copy(), 2 * N)
        ob = sfo.bh_overlay(nodes[:n], r.pos.copy(), r.mas.copy(), 3, M, np.full(sx * sy * 4, 0.25, np.float32), sx, sy)
        assert (rb != 0.25).sum() > 1000 and np.array_equal(rb, ob)
    r.close()


@pytest.mark.skipif(not sfo.ref_available("vis"), reason="reference visualization library not built")
@pytest.mark.parametrize("N,sx,sy,step", [(4096, 256, 256, 0), (20000, 320, 200, 7), (3000, 33, 17, 123), (9, 64, 64, 1),
                                          (30000, 60, 240, 4000), (65536, 1080, 1080, 2)])
def test_vis_draw_restatement_equals_reference(N, sx, sy, step):
    """The oracle's vis_draw() (mean radius, matrices, splat, finishing pass) against the unmodified reference's
    vis_init() + vis_draw() (visualization.c:37-739, oracle/_ref/libsfref_vis.so): matMVP, the integer accumulation
    buffer and the float render buffer bit for bit - images wider, narrower and much smaller than the projected ball."""
    r = sfo.RefSim("vis", N); r.initcond(); r.solv(0)
    mvp, ab, rb = r.vis_draw(step, sx, sy)
    rbar, mvp2, ab2, rb2 = sfo.vis_draw(r.pos.copy(), r.pot.copy(), step, sx, sy)
    r.close()
    assert np.array_equal(mvp, mvp2)
    assert np.array_equal(ab, ab2)
    assert np.array_equal(rb, rb2)
    assert ab[3::4].sum() > 0
