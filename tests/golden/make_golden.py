#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

The reference (Zi7ar21/starflood) has no tests and no golden files, so every fixture here is an
output of its own compiled code: oracle/build_ref.sh builds /root/reference/src into
oracle/_ref/libsfref_{pp,tree,pp_sph}.so and this script calls initcond_generate(), sim_solv() and
sim_step() in them through ctypes.  Run in the dev container (needs /root/reference):

    python tests/golden/make_golden.py

Fixtures (little-endian, numpy .npz, compressed):
  pp_n1024.npz      initcond_generate() state (pos, vel, mas) and solve_gravity_part_part() acc, pot
  bh_n4096.npz      root box of sim_solv(step 0), tree_build() node table (canonical numbering), BH acc, pot
  sph_n1024.npz     PP + solve_sph(): rad, rho, prs, acc (gravity + pressure), pot
  energy_pp_n512.npz  40 sim_step()s: per-step kinetic / potential totals recomputed from the ken / pen arrays
  vis_n4096.npz     vis_init() + vis_draw() of the reference's visualization.c (variant vis) on the BH state at step 5,
                    48 x 96 pixels (narrower than the projected ball): matMVP, integer accumulation buffer, render buffer
  pcg4d.npz         rng.h pcg4d on a few seeds (through initcond: first positions) - covered by pp_n1024
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sfo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    sfo.build()
    # ---- PP
    r = sfo.RefSim("pp", 1024); r.initcond(); r.solv(0)
    np.savez_compressed(os.path.join(OUT, "pp_n1024.npz"), pos=r.pos, vel=r.vel, mas=r.mas, acc=r.acc, pot=r.pot)
    # ---- BH
    r = sfo.RefSim("tree", 4096); r.initcond(); r.solv(0)
    nodes = r.nodes().copy()
    canon, n = sfo.tree_canonicalise(nodes)
    np.savez_compressed(os.path.join(OUT, "bh_n4096.npz"), pos=r.pos, mas=r.mas, bounds=np.array(nodes[0]["bounds"]),
                        nodes=canon[:n], acc=r.acc, pot=r.pot)
    # a later step number moves the jittered box (simulation.c:281-301)
    r.solv(7)
    nodes = r.nodes().copy()
    canon, n = sfo.tree_canonicalise(nodes)
    np.savez_compressed(os.path.join(OUT, "bh_n4096_step7.npz"), bounds=np.array(nodes[0]["bounds"]),
                        nodes=canon[:n], acc=r.acc, pot=r.pot)
    r.close()
    # ---- PP + SPH
    r = sfo.RefSim("pp_sph", 1024); r.initcond(); r.solv(0)
    np.savez_compressed(os.path.join(OUT, "sph_n1024.npz"), pos=r.pos, mas=r.mas, rad=r.rad, rho=r.rho, prs=r.prs,
                        acc=r.acc, pot=r.pot)
    # ---- energy series, PP N=512, 40 steps (main.c:396-412,423-550: initial solve, then sim_step per step)
    r = sfo.RefSim("pp", 512); r.initcond(); r.solv(0)
    T = []; V = []
    for s in range(40):
        r.step(s)
        T.append(float(np.sum(r.ken.astype(np.float64)))); V.append(float(np.sum(r.pen.astype(np.float64))))
    np.savez_compressed(os.path.join(OUT, "energy_pp_n512.npz"), T=np.array(T), V=np.array(V), pos=r.pos, vel=r.vel)
    # ---- vis_draw (visualization.c:253-739)
    r = sfo.RefSim("vis", 4096); r.initcond(); r.solv(0)
    mvp, ab, rb = r.vis_draw(5, 48, 96)
    np.savez_compressed(os.path.join(OUT, "vis_n4096.npz"), pos=r.pos, pot=r.pot, step=5, sizex=48, sizey=96, mvp=mvp,
                        atomic=ab, render=rb)
    r.close()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
