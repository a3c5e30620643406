"""Kick-drift-kick with engine forces against the reference's own time series (simulation.c:381-519)."""
import os

import numpy as np
import pytest

from oracle import sfo

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def energies(sim):
    """ken = 0.5 m v^2, pen = m pot in fp32 exactly as simulation.c:434-451, summed in double the same way
    tests/golden/make_*.py summed the reference's ken / pen arrays."""
    v = sim.vel.reshape(-1, 3)
    v2 = (v[:, 0] * v[:, 0]) + (v[:, 1] * v[:, 1]) + (v[:, 2] * v[:, 2])
    ken = np.float32(0.5) * sim.mas * v2
    pen = sim.mas * sim.pot
    return float(np.sum(ken.astype(np.float64))), float(np.sum(pen.astype(np.float64)))


def kdk(engine, sim, steps, tree, flavour=0):
    L = sfo.lib()
    N = sim.N
    T = np.zeros(steps); V = np.zeros(steps)
    sim.step_number = 0
    assert engine.sim_solv(sim, tree=tree, flavour=flavour) == 0
    for s in range(steps):
        sim.step_number = s
        L.sfo_kick(N, sim.vel, sim.acc, 1.0); L.sfo_drift(N, sim.pos, sim.vel, 1.0)
        assert engine.sim_solv(sim, tree=tree, flavour=flavour) == 0
        L.sfo_kick(N, sim.vel, sim.acc, 1.0)
        T[s], V[s] = energies(sim)
    return T, V


def test_pp_40_steps_matches_reference_series(engine):
    g = np.load(os.path.join(G, "energy_pp_n512.npz"))
    N = 512
    pos, vel, mas = sfo.initcond(N)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas
    T, V = kdk(engine, sim, 40, tree=False)
    E, Eref = T + V, g["T"] + g["V"]
    # N = 512 with eps = 1e-3 and dt = 1 is violently chaotic (E swings from -54 to -78 to -30 within 40 steps), so
    # round-off-level force differences grow: tight at the start, loose at the end (cf. SURVEY hard-part 6)
    d = np.abs(E - Eref) / abs(Eref[0])
    assert d[:5].max() <= 2e-6 and d.max() <= 1e-3


def test_bh_exact_walk_reproduces_reference_trajectory_bitwise(engine):
    """With the EXACT walk the whole sim_solv() is bit-identical to the reference, so a multi-step run
    (jittered root box per step, tree rebuild, walk) stays bit-identical to the oracle-driven run."""
    N = 4096
    pos, vel, mas = sfo.initcond(N)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas
    engine.tree_init(sim.tree, 8192)
    T, V = kdk(engine, sim, 8, tree=True)
    p, v = pos.copy(), vel.copy()
    L = sfo.lib()

    def solv(step):
        b = sfo.root_box(p, step)
        rc, nodes, n, *_ = sfo.tree_build_sorted(b, p, mas, 8192)
        return sfo.bh_walk(nodes[:n], p, mas)[:2]
    a, po = solv(0)
    for s in range(8):
        L.sfo_kick(N, v, a, 1.0); L.sfo_drift(N, p, v, 1.0)
        a, po = solv(s)
        L.sfo_kick(N, v, a, 1.0)
    assert np.array_equal(sim.pos, p) and np.array_equal(sim.vel, v) and np.array_equal(sim.pot, po)


@pytest.mark.parametrize("case", ["pp", "bh"])
def test_1000_step_energy_drift(engine, case):
    """BASELINE.json: total-energy drift over 1000 steps must match the reference's.  The reference series were
    produced by tests/golden/make_energy_series.py.  Tolerances follow SURVEY hard-part 6 / App. C.2: tight before
    the cold collapse, within the round-off sensitivity after it.  That sensitivity was measured on the engine
    itself (profiles/r01_energy_roundoff_sensitivity.txt, tools/energy_probe.py): perturbing ONE mass by 1 ulp moves
    max|dE|/|E0| anywhere between 2.2e-2 and 3.5e-2 (eight draws, peak around steps 50-85, just after the collapse),
    and the final drift by up to 5.5e-3; SURVEY measured 1.4e-2 for the reference with Kahan summation off.  The
    post-collapse bound is therefore 6e-2, and the final drift must agree to 3e-2."""
    name = {"pp": "energy1000_pp_n16384.npz", "bh": "energy1000_bh_n65536.npz"}[case]
    path = os.path.join(G, name)
    if not os.path.exists(path):
        pytest.skip("series fixture not generated")
    g = np.load(path)
    N = {"pp": 16384, "bh": 65536}[case]
    pos, vel, mas = sfo.initcond(N)
    sim = engine.Sim(N, pinned=True); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas
    if case == "bh":
        engine.tree_init(sim.tree, N)
    T, V = kdk(engine, sim, 1000, tree=(case == "bh"))
    E, Eref = T + V, g["T"] + g["V"]
    E0 = abs(Eref[0])
    d = np.abs(E - Eref) / E0
    if case == "bh":
        assert d.max() == 0.0            # exact walk: the series is the reference's, digit for digit
    else:
        assert d[:30].max() <= 1e-4 and d[:45].max() <= 1e-3 and d.max() <= 6e-2
        drift, drift_ref = (E[-1] - E[0]) / E0, (Eref[-1] - Eref[0]) / E0
        assert abs(drift - drift_ref) <= 3e-2
