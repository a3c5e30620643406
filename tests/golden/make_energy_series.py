#!/usr/bin/env python
"""1000-step total-energy series of the UNMODIFIED reference: PP N=16384 and BH N=65536,
defaults dt=1, eps=1e-3, G=1.  Follows main.c: initial sim_solv() at step_number 0, then for step s:
sim.step_number = s; sim_step().  Energies are the double sums of the ken / pen arrays the reference fills
(simulation.c:434-451).  Takes ~5 minutes on 8 cores; run in the dev container only."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sfo
OUT = os.path.dirname(os.path.abspath(__file__))
STEPS = 1000
for variant, N, name in (("pp", 16384, "energy1000_pp_n16384.npz"), ("tree", 65536, "energy1000_bh_n65536.npz")):
    t0 = time.time()
    r = sfo.RefSim(variant, N); r.initcond(); r.solv(0)
    T = np.zeros(STEPS); V = np.zeros(STEPS)
    for s in range(STEPS):
        r.step(s)
        T[s] = np.sum(r.ken.astype(np.float64)); V[s] = np.sum(r.pen.astype(np.float64))
    np.savez_compressed(os.path.join(OUT, name), T=T, V=V)
    print(name, "E0", T[0] + V[0], "E999", T[-1] + V[-1], "drift", (T[-1] + V[-1] - T[0] - V[0]) / abs(T[0] + V[0]), "s", time.time() - t0, flush=True)
