"""Scratch driver: PP + SPH parity and first timings on the GPU box."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import sfo
from starflood_b200 import engine as e

def rel_rms(a, ref):
    a = np.asarray(a, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.sqrt(np.sum((a - ref) ** 2) / np.sum(ref ** 2)))

e.init()
print(e.device_info())
for N in (1000, 4096, 65536):
    pos, vel, mas = sfo.initcond(N)
    sim = e.Sim(N); sim.pos[:] = pos; sim.mas[:] = mas
    rc = e.solve_gravity_part_part(sim)
    acc, pot = sfo.pp(pos, mas)
    a64, p64 = sfo.pp_f64(pos, mas)
    t = e.timings()
    print(N, 'rc', rc, 'acc relrms vs ref', rel_rms(sim.acc, acc), 'pot', rel_rms(sim.pot, pot),
          '| vs f64', rel_rms(sim.acc, a64), rel_rms(sim.pot, p64), '| ref vs f64', rel_rms(acc, a64),
          'ms', t.total_ms, t.solve_ms)
for T in (os.environ.get('SFE_PP_T', '4'),):
    for N in (65536, 262144, 1048576):
        pos, vel, mas = sfo.initcond(N)
        sim = e.Sim(N, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas
        for r in range(3):
            e.solve_gravity_part_part(sim); t = e.timings()
            print('T', T, 'N', N, 'total_ms', round(t.total_ms, 3), 'solve_ms', round(t.solve_ms, 3), 'h2d', round(t.h2d_ms, 3), 'd2h', round(t.d2h_ms, 3),
                  'inter/s %.3e' % (N * N / (t.solve_ms * 1e-3)), 'TFLOP/s(20)', round(20.0 * N * N / (t.solve_ms * 1e-3) / 1e12, 2))
# SPH
for N in (2048, 16384):
    pos, vel, mas, rad = sfo.initcond(N, sph=True)
    acc, pot = sfo.pp(pos, mas)
    rho = sfo.sph_density(pos, mas, rad); prs = sfo.sph_eos(rho)
    acc_ref = sfo.sph_force(pos, mas, rad, rho, prs, acc.copy())
    sim = e.Sim(N, sph=True); sim.pos[:] = pos; sim.mas[:] = mas; sim.rad[:] = rad
    e.solve_gravity_part_part(sim)
    g_only = sim.acc.copy()
    rc = e.solve_sph(sim); t = e.timings()
    print('SPH', N, rc, 'rho', rel_rms(sim.rho, rho), 'prs', rel_rms(sim.prs, prs), 'acc', rel_rms(sim.acc, acc_ref),
          'sph-only', rel_rms(sim.acc - g_only, acc_ref - acc), 'ms', t.solve_ms)
    # non-uniform h
    rad2 = (rad * (1.0 + 0.5 * np.sin(np.arange(N)))).astype(np.float32)
    rho = sfo.sph_density(pos, mas, rad2); prs = sfo.sph_eos(rho)
    acc_ref = sfo.sph_force(pos, mas, rad2, rho, prs, acc.copy())
    sim.rad[:] = rad2; sim.acc[:] = acc
    rc = e.solve_sph(sim)
    print('SPH varh', N, rc, 'rho', rel_rms(sim.rho, rho), 'prs', rel_rms(sim.prs, prs), 'acc', rel_rms(sim.acc, acc_ref))
N = 262144
pos, vel, mas, rad = sfo.initcond(N, sph=True)
sim = e.Sim(N, sph=True, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas; sim.rad[:] = rad
for r in range(2):
    e.solve_sph(sim); t = e.timings(); print('SPH N', N, 'solve_ms', t.solve_ms, 'total', t.total_ms)
