"""Scratch: how sensitive is the PP N=16384 1000-step energy series to a round-off-level perturbation?
Perturbs ONE particle's mass by 1 ulp (general-mass kernel) or nothing (equal-mass kernel) and prints the
deviation from the committed reference series in the three windows of tests/test_gpu_energy.py."""
import os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from oracle import sfo
from starflood_b200 import engine
import test_gpu_energy as t
engine.init(0)
g = np.load(os.path.join(t.G, "energy1000_pp_n16384.npz"))
N = 16384
Eref = g["T"] + g["V"]
for which in [None, 123, 4567, 9000, 15000, 1, 7777, 12001, 333]:
    pos, vel, mas = sfo.initcond(N)
    if which is not None:
        mas = mas.copy(); mas[which] = np.nextafter(mas[which], np.float32(1))
    sim = engine.Sim(N, pinned=True); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas
    T, V = t.kdk(engine, sim, 1000, tree=False)
    E = T + V
    d = np.abs(E - Eref) / abs(Eref[0])
    print("perturbed %-6s d[:30].max %.3e  d[:45].max %.3e  d.max %.3e (at %d)  d[100:].max %.3e  drift %.4e ref %.4e" % (
        which, d[:30].max(), d[:45].max(), d.max(), int(d.argmax()), d[100:].max(), (E[-1]-E[0])/abs(Eref[0]), (Eref[-1]-Eref[0])/abs(Eref[0])), flush=True)
