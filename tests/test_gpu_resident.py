"""Device-resident leapfrog (SURVEY 8f rows N1 + N2; sfe_sim_*, sfe_dev_kick_drift / root_box / kick_energy) against
the oracle's restatement of sim_step() / sim_solv() (simulation.c:231-519): vel, pos, ken, pen and the root box bit for
bit, the double sums T and V to double round-off (the reference folds them serially with Kahan compensation, the device
reduces them as exactly rounded double-double sums)."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_rms
from oracle import sfo

pytestmark = pytest.mark.gpu


def dev(t):
    return C.c_void_p(t.data_ptr())


def host_energies(vel, mas, pot):
    """ken / pen in fp32 as simulation.c:434-451; T, V with the oracle's Kahan folds (simulation.c:454-497)."""
    v = vel.reshape(-1, 3)
    v2 = (v[:, 0] * v[:, 0]) + (v[:, 1] * v[:, 1]) + (v[:, 2] * v[:, 2])
    ken = np.float32(0.5) * mas * v2
    pen = mas * pot
    T, V = sfo.energies(vel, mas, pot)
    return ken, pen, T, V


@pytest.mark.parametrize("N", [1, 9, 4096, 100003, 1 << 20])
def test_root_box_on_device_is_the_reference_box(engine, N):
    """simulation.c:245-301 for several step numbers (different pcg4d jitters) and sizes; N = 2^20 is C2's box."""
    import torch
    L = engine.lib()
    pos, vel, mas = sfo.initcond(N)
    device = torch.device("cuda", 0)
    pm4 = torch.empty(4 * N, dtype=torch.float32, device=device)
    s = torch.cuda.current_stream().cuda_stream
    assert L.sfe_dev_pack_pm4(dev(torch.from_numpy(pos).to(device)), dev(torch.from_numpy(mas).to(device)), dev(pm4), N, s) == 0
    for step in (0, 1, 2, 77, 12345, 0xFFFFFFFF):
        box = np.zeros(6, np.float32)
        assert L.sfe_dev_root_box(dev(pm4), N, step, box.ctypes.data, s) == 0, engine.last_error()
        assert box.tobytes() == sfo.root_box(pos, step).tobytes(), (N, step)
    if N == 1 << 20:
        # SURVEY 8(d) quotes C2's box at step 0
        assert np.allclose(sfo.root_box(pos, 0), [-1.37745, -1.27724, -1.71780, 1.62190, 1.72211, 1.28155], atol=1e-5)


def test_root_box_of_scaled_snapshots(engine):
    """Many different sums (scaled / shifted copies of one snapshot): the exactly rounded device sum and the
    reference's Kahan fold must round to the same fp32 box every time."""
    import torch
    L = engine.lib()
    N = 20011
    pos, vel, mas = sfo.initcond(N)
    device = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(5)
    dm = torch.from_numpy(mas).to(device)
    pm4 = torch.empty(4 * N, dtype=torch.float32, device=device)
    for k in range(40):
        q = (pos * np.float32(rng.uniform(0.01, 50.0)) + np.float32(rng.normal(0, 0.3))).astype(np.float32)
        assert L.sfe_dev_pack_pm4(dev(torch.from_numpy(q).to(device)), dev(dm), dev(pm4), N, s) == 0
        box = np.zeros(6, np.float32)
        assert L.sfe_dev_root_box(dev(pm4), N, k, box.ctypes.data, s) == 0
        assert box.tobytes() == sfo.root_box(q, k).tobytes(), k


def test_kick_drift_and_kick_energy_slices_bitwise(engine):
    """The O(N) kernels on two target slices (the multi-GPU partition) reproduce the host loops bit for bit."""
    import torch
    L = engine.lib()
    N = 30007
    rng = np.random.default_rng(11)
    pos, vel, mas = sfo.initcond(N)
    mas = (mas * rng.uniform(0.5, 2.0, N)).astype(np.float32)
    acc = rng.normal(0, 1e-3, 3 * N).astype(np.float32)
    pot = -np.abs(rng.normal(1, 0.3, N)).astype(np.float32)
    dt = 0.75
    device = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    pm4 = torch.empty(4 * N, dtype=torch.float32, device=device)
    assert L.sfe_dev_pack_pm4(dev(torch.from_numpy(pos).to(device)), dev(torch.from_numpy(mas).to(device)), dev(pm4), N, s) == 0
    dvel = torch.from_numpy(vel).to(device); dacc = torch.from_numpy(acc).to(device); dpot = torch.from_numpy(pot).to(device)
    ken = torch.zeros(N, dtype=torch.float32, device=device); pen = torch.zeros(N, dtype=torch.float32, device=device)
    tv = torch.zeros(4, dtype=torch.float64, device=device)
    cuts = [(0, 12000), (12000, N - 12000)]
    for i0, ni in cuts:
        assert L.sfe_dev_kick_drift(dev(pm4), dev(dvel), dev(dacc), i0, ni, dt, s) == 0
    # host loops (simulation.c:381-395)
    SL = sfo.lib()
    v, p = vel.copy(), pos.copy()
    SL.sfo_kick(N, v, acc, dt); SL.sfo_drift(N, p, v, dt)
    torch.cuda.synchronize()
    got = pm4.cpu().numpy().reshape(-1, 4)
    assert np.array_equal(got[:, :3].ravel(), p) and np.array_equal(got[:, 3], mas)
    assert np.array_equal(dvel.cpu().numpy(), v)
    # kick_1 + energies by slice (simulation.c:423-504)
    for k, (i0, ni) in enumerate(cuts):
        assert L.sfe_dev_kick_energy(dev(pm4), dev(dvel), dev(dacc), dev(dpot), dev(ken), dev(pen), i0, ni, dt, 1,
                                     C.c_void_p(tv.data_ptr() + 16 * k), s) == 0
    SL.sfo_kick(N, v, acc, dt)
    hk, hp, T, V = host_energies(v, mas, pot)
    torch.cuda.synchronize()
    assert np.array_equal(dvel.cpu().numpy(), v)
    assert np.array_equal(ken.cpu().numpy(), hk) and np.array_equal(pen.cpu().numpy(), hp)
    t = tv.cpu().numpy()
    assert abs((t[0] + t[2]) - T) <= 4e-16 * abs(T) and abs((t[1] + t[3]) - V) <= 4e-16 * abs(V)
    # energies only (do_kick = 0) leave vel alone
    assert L.sfe_dev_kick_energy(dev(pm4), dev(dvel), dev(dacc), dev(dpot), dev(ken), dev(pen), 0, N, dt, 0, dev(tv), s) == 0
    torch.cuda.synchronize()
    assert np.array_equal(dvel.cpu().numpy(), v)
    t = tv.cpu().numpy()
    assert abs(t[0] - T) <= 4e-16 * abs(T) and abs(t[1] - V) <= 4e-16 * abs(V)


def oracle_run(pos, vel, mas, steps, nodes_max, dt=1.0):
    """main.c's loop with ENABLE_TREE: initial sim_solv(), then sim_step() per step (step_number = s)."""
    L = sfo.lib()
    N = len(mas)
    p, v = pos.copy(), vel.copy()

    def solv(step):
        b = sfo.root_box(p, step)
        rc, nodes, n, *_ = sfo.tree_build_sorted(b, p, mas, nodes_max)
        a, po, _ = sfo.bh_walk(nodes[:n], p, mas)
        return a, po, b
    a, po, b = solv(0)
    E = np.zeros((steps, 3))
    for s in range(steps):
        L.sfo_kick(N, v, a, dt); L.sfo_drift(N, p, v, dt)
        a, po, b = solv(s)
        L.sfo_kick(N, v, a, dt)
        ken, pen, T, V = host_energies(v, mas, po)
        E[s] = (V + T, V, T)
    return p, v, a, po, ken, pen, E, b


def test_resident_bh_run_is_the_reference_trajectory(engine):
    """Upload once, sfe_sim_solv() + 6 resident steps (jittered box, build, EXACT walk, kicks, energies on the device),
    download once: every block of sim.mem equals the oracle-driven run bit for bit."""
    N = 6000
    pos, vel, mas = sfo.initcond(N)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas
    prm = engine.sim_params(tree=True, nodes_max=2 * N)
    assert engine.sim_upload(sim) == 0, engine.last_error()
    sim.step_number = 0
    assert engine.sim_solv_resident(sim, prm) == 0, engine.last_error()
    rc, E = engine.sim_step(sim, prm, nsteps=6)
    assert rc == 0, engine.last_error()
    sim.mem[:] = -7.0
    assert engine.sim_download(sim) == 0
    p, v, a, po, ken, pen, Eo, b = oracle_run(pos, vel, mas, 6, 2 * N)
    assert np.array_equal(sim.pos, p) and np.array_equal(sim.vel, v) and np.array_equal(sim.mas, mas)
    assert np.array_equal(sim.acc, a) and np.array_equal(sim.pot, po)
    assert np.array_equal(sim.ken, ken) and np.array_equal(sim.pen, pen)
    assert np.array_equal(engine.sim_last_box(), b)
    assert np.all(np.abs(E - Eo) <= 1e-15 * np.abs(Eo).max())
    # one more step through sfe_sim_step() continues the same trajectory (step numbers matter: the box jitter)
    sim.step_number = 6
    rc, E1 = engine.sim_step(sim, prm)
    assert rc == 0
    assert engine.sim_download(sim, ["pos"]) == 0
    p7 = oracle_run(pos, vel, mas, 7, 2 * N)[0]
    assert np.array_equal(sim.pos, p7)


def test_resident_pp_sph_run_matches_host_path(engine):
    """PP + SPH steps on the resident state against the same steps driven through the host entry points."""
    N = 3000
    pos, vel, mas, rad = sfo.initcond(N, outflow=0.0, sph=True)
    a = engine.Sim(N, sph=True); a.pos[:] = pos; a.vel[:] = vel; a.mas[:] = mas; a.rad[:] = rad
    prm = engine.sim_params(tree=False, sph=True)
    assert engine.sim_upload(a) == 0
    assert engine.sim_solv_resident(a, prm) == 0, engine.last_error()
    rc, E = engine.sim_step(a, prm, nsteps=3)
    assert rc == 0, engine.last_error()
    assert engine.sim_download(a) == 0
    b = engine.Sim(N, sph=True); b.pos[:] = pos; b.vel[:] = vel; b.mas[:] = mas; b.rad[:] = rad
    L = sfo.lib()
    assert engine.sim_solv(b, tree=False, sph=True) == 0
    for s in range(3):
        L.sfo_kick(N, b.vel, b.acc, 1.0); L.sfo_drift(N, b.pos, b.vel, 1.0)
        assert engine.sim_solv(b, tree=False, sph=True) == 0
        L.sfo_kick(N, b.vel, b.acc, 1.0)
    # same kernels on the same inputs: identical
    for name in ("pos", "vel", "acc", "pot", "rho", "prs", "rad"):
        assert np.array_equal(a.find(name), b.find(name)), name
    ken, pen, T, V = host_energies(b.vel, b.mas, b.pot)
    assert np.array_equal(a.ken, ken) and np.array_equal(a.pen, pen)
    assert abs(E[-1, 2] - T) <= 1e-15 * abs(T) and abs(E[-1, 1] - V) <= 1e-15 * abs(V)


def test_async_raw_snapshots_are_the_state_block(engine, tmp_path):
    """sfe_sim_save_raw_async() writes sim_save_raw()'s file (real[N*SIM_DOF], simulation_io.c:126-133)
    from the resident state while the leapfrog goes on.  Every file must equal the state block that a second,
    step-by-step run downloads at the same step."""
    N, steps = 20000, 5
    for sph in (False, True):
        if sph:
            pos, vel, mas, rad = sfo.initcond(N, outflow=0.0, sph=True)
        else:
            pos, vel, mas = sfo.initcond(N)
        prm = engine.sim_params(tree=not sph, sph=sph, nodes_max=2 * N)

        def fresh():
            sim = engine.Sim(N, sph=sph); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas
            if sph:
                sim.rad[:] = rad
            assert engine.sim_upload(sim) == 0 and engine.sim_solv_resident(sim, prm) == 0, engine.last_error()
            return sim
        sim = fresh()
        names = [str(tmp_path / ("sph%d_step_%04d.raw" % (sph, k))) for k in range(steps + 1)]
        assert engine.sim_save_raw_async(names[0]) == 0, engine.last_error()
        for k in range(steps):                       # no waiting between steps: up to two snapshots are in flight
            sim.step_number = k
            assert engine.sim_step(sim, prm)[0] == 0, engine.last_error()
            assert engine.sim_save_raw_async(names[k + 1]) == 0, engine.last_error()
        assert engine.sim_save_wait() == 0, engine.last_error()
        ref = fresh()
        for k in range(steps + 1):
            if k:
                ref.step_number = k - 1
                assert engine.sim_step(ref, prm)[0] == 0
            assert engine.sim_download(ref) == 0
            data = np.fromfile(names[k], np.float32)
            assert data.size == N * ref.dof and data.tobytes() == ref.mem.tobytes(), (sph, k)
        # overwriting an existing, longer file leaves exactly the block
        with open(names[0], "ab") as f:
            f.write(b"x" * 1000)
        assert engine.sim_save_raw_async(names[0]) == 0 and engine.sim_save_wait() == 0
        assert np.fromfile(names[0], np.float32).tobytes() == ref.mem.tobytes()
    # writer errors surface in sfe_sim_save_wait(), once
    assert engine.sim_save_raw_async(str(tmp_path / "no_such_dir" / "x.raw")) == 0
    assert engine.sim_save_wait() == engine.STARFLOOD_FAILURE and "open" in engine.last_error()
    assert engine.sim_save_wait() == 0
    assert engine.lib().sfe_sim_save_raw_async(None) == engine.STARFLOOD_FAILURE
    engine.lib().sfe_sim_release()
    assert engine.sim_save_raw_async(str(tmp_path / "y.raw")) == engine.STARFLOOD_FAILURE      # nothing resident


def test_resident_errors(engine):
    L = engine.lib()
    engine.lib().sfe_sim_release()
    sim = engine.Sim(64)
    prm = engine.sim_params()
    assert engine.sim_solv_resident(sim, prm) == engine.STARFLOOD_FAILURE and "sfe_sim_upload" in engine.last_error()
    assert L.sfe_sim_upload(sim.mem.ctypes.data, 64, 12) == engine.STARFLOOD_FAILURE      # SIM_DOF is 13 or 16
    sim.mas[:] = 1.0; sim.pos[:] = np.linspace(-1, 1, 192)
    assert engine.sim_upload(sim) == 0
    other = engine.Sim(65)
    assert engine.sim_download(other) == engine.STARFLOOD_FAILURE                          # size mismatch
    assert engine.sim_step(sim, engine.sim_params(tree=False, sph=True))[0] == engine.STARFLOOD_FAILURE   # no SPH blocks
    # node overflow is raised, not ignored (the reference's sim_solv() ignores it, simulation.c:312)
    assert engine.sim_solv_resident(sim, engine.sim_params(nodes_max=4)) == engine.STARFLOOD_FAILURE
    assert engine.sim_solv_resident(sim, engine.sim_params(nodes_max=1024)) == 0
