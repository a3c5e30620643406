"""Scratch driver: SPH parity + timing (cell list vs all pairs: run again with SFE_SPH_ALL_PAIRS=1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import sfo
from starflood_b200 import engine as e

def rel_rms(a, ref):
    a = np.asarray(a, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.sqrt(np.sum((a - ref) ** 2) / np.sum(ref ** 2)))

e.init()
tag = "all-pairs" if os.environ.get("SFE_SPH_ALL_PAIRS") else "cells"
for N in (20000, 40000):
    pos, vel, mas, rad = sfo.initcond(N, sph=True)
    acc0 = np.zeros(3 * N, np.float32)
    rho = sfo.sph_density(pos, mas, rad); prs = sfo.sph_eos(rho)
    acc = sfo.sph_force(pos, mas, rad, rho, prs, acc0.copy())
    sim = e.Sim(N, sph=True); sim.pos[:] = pos; sim.mas[:] = mas; sim.rad[:] = rad; sim.acc[:] = acc0
    rc = e.solve_sph(sim); t = e.timings()
    print(tag, N, rc, 'rho', rel_rms(sim.rho, rho), 'prs', rel_rms(sim.prs, prs), 'acc', rel_rms(sim.acc, acc), 'solve_ms', t.solve_ms, flush=True)
    rad2 = (rad * (1.0 + 0.3 * np.sin(np.arange(N)))).astype(np.float32)
    rho = sfo.sph_density(pos, mas, rad2); prs = sfo.sph_eos(rho)
    acc = sfo.sph_force(pos, mas, rad2, rho, prs, acc0.copy())
    sim.rad[:] = rad2; sim.acc[:] = acc0
    rc = e.solve_sph(sim); t = e.timings()
    print(tag, 'varh', N, rc, 'rho', rel_rms(sim.rho, rho), 'prs', rel_rms(sim.prs, prs), 'acc', rel_rms(sim.acc, acc), 'solve_ms', t.solve_ms, flush=True)
N = 262144
pos, vel, mas, rad = sfo.initcond(N, outflow=0.0, sph=True)
sim = e.Sim(N, sph=True, pinned=True); sim.pos[:] = pos; sim.mas[:] = mas; sim.rad[:] = rad
for r in range(3):
    sim.acc[:] = 0
    e.solve_sph(sim); t = e.timings(); print(tag, 'SPH N', N, 'solve_ms', t.solve_ms, 'total', t.total_ms, flush=True)
i0, i1 = 100000, 100256
rho_s = sfo.sph_density(pos, mas, rad, i0, i1)
acc_s = sfo.sph_force(pos, mas, rad, sim.rho, sim.prs, np.zeros(3 * N, np.float32), i0, i1)
print(tag, 'C3 sample rho', rel_rms(sim.rho[i0:i1], rho_s[i0:i1]), 'acc', rel_rms(sim.acc[3*i0:3*i1], acc_s[3*i0:3*i1]))
