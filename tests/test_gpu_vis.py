"""vis_draw() on the device (second half; visualization.c:253-739) through sfe_vis_draw() /
sfe_sim_vis_draw(): the INTEGER accumulation buffer must equal the oracle's restatement bit for bit (integer adds
commute, and csrc/sfe_libm.cuh reproduces the host libm's logf / sinf / cosf bit patterns), matMVP bit for bit, the float
render buffer to within one fp32 ulp (the tanh tonemap is evaluated in double by a different library and rounded once).
The oracle itself is pinned against the unmodified reference's visualization.c (tests/test_oracle_vs_ref.py) and, when the
reference library is present, the engine is compared with it directly as well."""
import numpy as np
import pytest

from oracle import sfo

pytestmark = pytest.mark.gpu


def ulp_diff(a, b):
    ia = a.view(np.int32).astype(np.int64); ib = b.view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


def state(engine, N, step, pot_scale=1.0):
    pos, vel, mas = sfo.initcond(N)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas
    # a plausible potential without running a solver: -(3/2 - r^2/2) of the uniform unit ball, times a scale
    r2 = (pos.reshape(-1, 3).astype(np.float64) ** 2).sum(1)
    sim.pot[:] = (-(1.5 - 0.5 * r2) * 4.18879 * pot_scale).astype(np.float32)
    sim.step_number = step
    return sim


def draw(engine, sim, sx, sy):
    vis = engine.Vis(); engine.vis_init(vis, sx, sy)
    assert engine.vis_draw(vis, sim) == 0, engine.last_error()
    return vis


@pytest.mark.parametrize("N,sx,sy,step", [(20000, 320, 200, 0), (4096, 256, 256, 7), (3000, 33, 17, 123), (9, 64, 64, 1),
                                          (1, 16, 16, 0), (50000, 100, 400, 4000)])
def test_vis_draw_equals_the_oracle(engine, N, sx, sy, step):
    """Includes images narrower than the projected ball (samples fall outside and are dropped, coordinates in (-1, 0)
    truncate to pixel 0 like the reference's (int) cast) and tiny images (every window straddles the border)."""
    sim = state(engine, N, step)
    vis = draw(engine, sim, sx, sy)
    rbar, mvp, ab, rb = sfo.vis_draw(sim.pos.copy(), sim.pot.copy(), step, sx, sy)
    assert np.array_equal(vis.matMVP, mvp)
    assert abs(vis.average_radius - rbar) <= 4e-16 * rbar
    assert np.array_equal(vis.atomic_buffer, ab), "accumulation buffer differs in %d entries" % (vis.atomic_buffer != ab).sum()
    assert ab[3::4].sum() > 0
    assert ulp_diff(vis.render_buffer, rb).max() <= 1
    assert np.array_equal(vis.render_buffer[3::4], np.ones(sx * sy, np.float32))


def test_vis_draw_extreme_inputs(engine):
    """Potentials large enough to leave the fast path of cosf (|x| >= 120: the documented fall-back, not bit-pinned, so
    only particles with small potentials are compared exactly), a particle far outside the image, NaN / inf positions."""
    N, sx, sy = 4096, 128, 128
    sim = state(engine, N, 3)
    sim.pos[0:3] = [1.0e6, -1.0e6, 0.0]          # far outside
    sim.pos[3:6] = [np.nan, 0.0, 0.0]            # the reference's (int)NaN is INT_MIN: dropped
    sim.pos[6:9] = [np.inf, 0.0, 0.0]
    vis = draw(engine, sim, sx, sy)
    if np.isfinite(vis.average_radius):          # mean radius is NaN here (a NaN position): the matrix is NaN, nothing lands
        pytest.fail("NaN position must poison the mean radius like the reference's sum")
    rbar, mvp, ab, rb = sfo.vis_draw(sim.pos.copy(), sim.pot.copy(), 3, sx, sy)
    assert np.array_equal(vis.atomic_buffer, ab) and not ab.any()
    # without the NaN / inf: the far particle is simply dropped
    sim.pos[3:9] = 0.25
    vis = draw(engine, sim, sx, sy)
    rbar, mvp, ab, rb = sfo.vis_draw(sim.pos.copy(), sim.pot.copy(), 3, sx, sy)
    assert np.array_equal(vis.matMVP, mvp) and np.array_equal(vis.atomic_buffer, ab)
    # bad arguments fail loudly
    bad = engine.Vis(); engine.vis_init(bad, 0, 16)
    assert engine.vis_draw(bad, sim) == engine.STARFLOOD_FAILURE and "not supported" in engine.last_error()
    empty = engine.Sim(0)
    ok = engine.Vis(); engine.vis_init(ok, 16, 16)
    assert engine.vis_draw(ok, empty) == engine.STARFLOOD_FAILURE


def test_vis_draw_at_workload_size(engine):
    """C2 size (N = 2^20, 1.3e8 samples) at the reference's default 1080 x 1080 image, on a real Barnes-Hut potential."""
    N, sx, sy = 1 << 20, 1080, 1080
    pos, vel, mas = sfo.initcond(N)
    sim = engine.Sim(N, pinned=True); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas
    engine.tree_init(sim.tree, N)
    assert engine.sim_solv(sim, tree=True) == 0, engine.last_error()
    vis = draw(engine, sim, sx, sy)
    t = engine.timings()
    rbar, mvp, ab, rb = sfo.vis_draw(sim.pos.copy(), sim.pot.copy(), 0, sx, sy)
    assert np.array_equal(vis.matMVP, mvp)
    assert np.array_equal(vis.atomic_buffer, ab), "accumulation buffer differs in %d entries" % (vis.atomic_buffer != ab).sum()
    assert int(ab[3::4].sum()) == N * 128                      # every sample lands inside this image
    d = ulp_diff(vis.render_buffer, rb)
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    print("vis_draw N=2^20 1080x1080: solve %.3f ms, total %.3f ms" % (t.solve_ms, t.total_ms))


def test_resident_vis_draw(engine):
    """sfe_sim_vis_draw(): the frame of the HBM-resident state after two leapfrog steps equals vis_draw() of the
    downloaded state."""
    N, sx, sy = 30000, 300, 300
    pos, vel, mas = sfo.initcond(N)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas
    params = engine.sim_params(tree=True, nodes_max=2 * N)
    engine.sim_upload(sim)
    engine.sim_solv_resident(sim, params)
    engine.sim_step(sim, params, nsteps=2)
    vis = engine.Vis(); engine.vis_init(vis, sx, sy)
    assert engine.sim_vis_draw(vis, 2) == 0, engine.last_error()
    engine.sim_download(sim)
    rbar, mvp, ab, rb = sfo.vis_draw(sim.pos.copy(), sim.pot.copy(), 2, sx, sy)
    assert np.array_equal(vis.matMVP, mvp) and np.array_equal(vis.atomic_buffer, ab)
    assert ulp_diff(vis.render_buffer, rb).max() <= 1
    engine.lib().sfe_sim_release()


@pytest.mark.skipif(not sfo.ref_available("vis"), reason="reference visualization library not built")
def test_vis_draw_equals_the_reference(engine):
    """Directly against the unmodified reference's vis_init() + vis_draw() (oracle/_ref/libsfref_vis.so)."""
    N, sx, sy, step = 20000, 400, 300, 11
    ref = sfo.RefSim("vis", N); ref.initcond(); ref.solv(0)
    mvp, ab, rb = ref.vis_draw(step, sx, sy)
    sim = engine.Sim(N); sim.mem[:] = ref.mem; sim.step_number = step
    vis = draw(engine, sim, sx, sy)
    ref.close()
    assert np.array_equal(vis.matMVP, mvp)
    assert np.array_equal(vis.atomic_buffer, ab)
    assert ulp_diff(vis.render_buffer, rb).max() <= 1


@pytest.mark.parametrize("samples,exposure,fmt", [(5, None, 0), (100, -3.0, 2), (33, 0.5, 1), (1, -1.0, 0)])
def test_vis_draw_other_configurations(engine, samples, exposure, fmt, monkeypatch):
    """SPATIAL_SAMPLES that are not a multiple of the warp size, other EXPOSURE values, VIS_FILE_FORMAT 0 (no tonemap: the
    render buffer is then a double product rounded once and must be bit-identical), AUTOSCALING / ORTHO_SCALE off."""
    N, sx, sy, step = 6000, 96, 160, 9
    sim = state(engine, N, step)
    monkeypatch.setattr(engine, "SPATIAL_SAMPLES", samples)
    monkeypatch.setattr(engine, "EXPOSURE", exposure)
    monkeypatch.setattr(engine, "VIS_FILE_FORMAT", fmt)
    auto = samples != 33
    monkeypatch.setattr(engine, "AUTOSCALING", auto)
    monkeypatch.setattr(engine, "ORTHO_SCALE", None if samples == 1 else -1.0)
    vis = draw(engine, sim, sx, sy)
    rbar, mvp, ab, rb = sfo.vis_draw(sim.pos.copy(), sim.pot.copy(), step, sx, sy, samples=samples, exposure=exposure,
                                     tonemap=1 if fmt > 0 else 0, ortho_exp2=0.0 if samples == 1 else -1.0, autoscale=1 if auto else 0)
    assert np.array_equal(vis.matMVP, mvp)
    assert np.array_equal(vis.atomic_buffer, ab)
    assert int(ab[3::4].sum()) > 0
    if fmt == 0:
        assert np.array_equal(vis.render_buffer, rb)
    else:
        assert ulp_diff(vis.render_buffer, rb).max() <= 1
