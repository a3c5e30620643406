"""The device entry points (sfe_dev_*, include/sfengine.h) that the multi-GPU flow of SURVEY 8(e) uses: target
slices of PP / BH / SPH computed on device-resident {x,y,z,m} records must tile the answer of the host entry
points (BH EXACT: bit for bit; PP / SPH: the grouping of partial sums depends on the slice, so to round-off)."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_rms
from oracle import sfo

pytestmark = pytest.mark.gpu


def dev(t):
    return C.c_void_p(t.data_ptr())


@pytest.fixture(scope="module")
def torch_dev():
    import torch
    return torch, torch.device("cuda", 0)


def upload(torch, device, pos, mas, L, N):
    dpos = torch.from_numpy(pos).to(device); dmas = torch.from_numpy(mas).to(device)
    pm4 = torch.empty(4 * N, dtype=torch.float32, device=device)
    s = torch.cuda.current_stream().cuda_stream
    assert L.sfe_dev_pack_pm4(dev(dpos), dev(dmas), dev(pm4), N, s) == 0
    return pm4, s


@pytest.mark.parametrize("N,parts", [(3000, 3), (40000, 4)])
def test_sph_slices_tile_the_host_answer(engine, torch_dev, N, parts):
    """N = 40000 with h = 0.1 takes the cell-list path (sfe_sph_cells.cu), N = 3000 the all-pairs kernels."""
    torch, device = torch_dev
    L = engine.lib()
    pos, vel, mas, rad = sfo.initcond(N, outflow=0.0, sph=True)
    acc0 = np.linspace(-1, 1, 3 * N).astype(np.float32)
    sim = engine.Sim(N, sph=True); sim.pos[:] = pos; sim.mas[:] = mas; sim.rad[:] = rad; sim.acc[:] = acc0
    assert engine.solve_sph(sim) == 0, engine.last_error()

    pm4, s = upload(torch, device, pos, mas, L, N)
    drad = torch.from_numpy(rad).to(device)
    rho = torch.full((N,), -1.0, dtype=torch.float32, device=device)
    prs = torch.empty(N, dtype=torch.float32, device=device)
    acc = torch.from_numpy(acc0).to(device)
    cuts = [(N * k) // parts for k in range(parts + 1)]
    for k in range(parts):
        i0, ni = cuts[k], cuts[k + 1] - cuts[k]
        assert L.sfe_dev_sph_density(dev(pm4), dev(drad), N, i0, ni, dev(rho), s) == 0, engine.last_error()
    assert L.sfe_dev_sph_eos(dev(rho), dev(prs), N, engine.EOS_K, engine.EOS_N, s) == 0
    for k in range(parts):
        i0, ni = cuts[k], cuts[k + 1] - cuts[k]
        assert L.sfe_dev_sph_force(dev(pm4), dev(drad), dev(rho), dev(prs), N, i0, ni, engine.SPH_CUTOFF2, dev(acc), s) == 0
    torch.cuda.synchronize()
    assert (rho.cpu().numpy() > 0).all()                      # every target was written by exactly one slice
    assert rel_rms(rho.cpu().numpy(), sim.rho) <= 1e-6
    assert rel_rms(prs.cpu().numpy(), sim.prs) <= 2e-6
    assert rel_rms(acc.cpu().numpy() - acc0, sim.acc - acc0) <= 2e-6


def test_pp_and_bh_slices_on_device(engine, torch_dev):
    torch, device = torch_dev
    L = engine.lib()
    N = 20000
    pos, vel, mas = sfo.initcond(N)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.mas[:] = mas
    pm4, s = upload(torch, device, pos, mas, L, N)
    acc = torch.zeros(3 * N, dtype=torch.float32, device=device); pot = torch.zeros(N, dtype=torch.float32, device=device)
    # PP
    assert engine.solve_gravity_part_part(sim) == 0
    for i0, ni in ((0, 7000), (7000, 13000)):
        assert L.sfe_dev_pp(dev(pm4), N, i0, ni, dev(acc), dev(pot), engine.G, engine.EPS2, s) == 0
    torch.cuda.synchronize()
    assert rel_rms(acc.cpu().numpy(), sim.acc) <= 1e-6 and rel_rms(pot.cpu().numpy(), sim.pot) <= 1e-6
    # BH: full tree on the device, walk by slices
    engine.tree_init(sim.tree, N)
    assert engine.sim_solv(sim, tree=True) == 0
    box = np.ascontiguousarray(engine.root_box(sim), np.float32)
    assert L.sfe_dev_tree_build(box.ctypes.data, dev(pm4), N, engine.TREE_DEPTH_MAX, N, s) == 0
    for i0, ni in ((0, 5000), (5000, 15000)):
        assert L.sfe_dev_bh(dev(pm4), N, i0, ni, dev(acc), dev(pot), engine.G, engine.EPS2, engine.TREE_THETA, engine.WALK_EXACT, s) == 0
    torch.cuda.synchronize()
    assert np.array_equal(acc.cpu().numpy(), sim.acc) and np.array_equal(pot.cpu().numpy(), sim.pot)
