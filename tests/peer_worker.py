"""One rank of the multi-rank tests (tests/test_gpu_peer.py starts `world` of these).

    python tests/peer_worker.py <rank> <world> <port>

Ranks map to GPUs round-robin; on a single-GPU box all ranks share device 0 (CUDA IPC works between processes on one
device, the exchanges then wait for the driver's time slicing - slow but the same code path).  The host-side transport
for the 64-byte window handles is torch.distributed over gloo.  Prints PEER_OK on success."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class _Window:
    """Device memory of the exchange window as a torch tensor (via __cuda_array_interface__)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    import torch
    import torch.distributed as dist
    from oracle import sfo
    from starflood_b200 import engine as e
    from test_gpu_resident import oracle_run

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ndev = torch.cuda.device_count()
    device = rank % ndev
    torch.cuda.set_device(device)
    e.init(device)
    L = e.lib()
    s = torch.cuda.current_stream().cuda_stream

    def allgather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    # ---- 1. raw exchange: ragged slices (odd byte counts, unaligned offsets), several rounds, ping-pong regions
    half = 1 << 20
    ptr = e.peer_setup(rank, world, 2 * half, allgather)
    win = torch.as_tensor(_Window(ptr, 2 * half), device="cuda")
    cuts = [(half * r) // world + (3 * r if r else 0) for r in range(world)] + [half - 5]
    for rnd in range(6):
        reg = (rnd & 1) * half
        a, b = cuts[rank], cuts[rank + 1]
        win[reg + a:reg + b] = ((torch.arange(a, b, device="cuda") * 7 + rank * 31 + rnd) % 251).to(torch.uint8)
        assert L.sfe_dev_peer_put(reg + a, b - a, s) == 0, e.last_error()
        got = win[reg + cuts[0]:reg + cuts[-1]].cpu().numpy()
        for r in range(world):
            x = np.arange(cuts[r], cuts[r + 1])
            want = ((x * 7 + r * 31 + rnd) % 251).astype(np.uint8)
            assert np.array_equal(got[cuts[r] - cuts[0]:cuts[r + 1] - cuts[0]], want), (rank, rnd, r)
    assert L.sfe_dev_peer_barrier(s) == 0
    torch.cuda.synchronize()
    # a range outside the window is refused
    assert L.sfe_dev_peer_put(2 * half - 8, 64, s) == e.STARFLOOD_FAILURE
    del win
    assert e.peer_close(dist.barrier) == 0, e.last_error()

    # ---- 1b. routed walk (sfe_dev_bh_peer): slot ranges walked, results delivered to the owners of the particle indices
    N = 30011
    pos, vel, mas = sfo.initcond(N)
    pos[3 * 5:3 * 5 + 3] = [30.0, -20.0, 25.0]                       # one particle outside the root box
    box = sfo.root_box(pos, 0)
    ptr = e.peer_setup(rank, world, 32 * N, allgather)
    win = torch.as_tensor(_Window(ptr, 32 * N), device="cuda").view(torch.float32)
    pm = np.empty((N, 4), np.float32); pm[:, :3] = pos.reshape(-1, 3); pm[:, 3] = mas
    win[:4 * N] = torch.from_numpy(pm.ravel()).cuda()
    win[4 * N:] = -7.0
    ref = e.Sim(N); ref.pos[:] = pos; ref.mas[:] = mas
    e.tree_init(ref.tree, 2 * N); ref.tree.bounds[:] = box
    for flavour in (e.WALK_EXACT, e.WALK_FAST):
        assert e.tree_build(ref.tree, N, ref.pos, ref.mas) == 0 and e.solve_gravity_part_tree(ref, flavour) == 0
        assert L.sfe_dev_tree_build(box.ctypes.data, win.data_ptr(), N, e.TREE_DEPTH_MAX, 2 * N, s) == 0, e.last_error()
        assert L.sfe_dev_bh_peer(win.data_ptr(), N, 16 * N, 28 * N, e.G, e.EPS2, e.TREE_THETA, flavour, s) == 0, e.last_error()
        got = win[4 * N:].cpu().numpy()
        i0, i1 = e.sim_slice(N, rank, world)
        assert np.array_equal(got[3 * i0:3 * i1], ref.acc[3 * i0:3 * i1]) and np.array_equal(got[3 * N + i0:3 * N + i1], ref.pot[i0:i1])
        rest = np.concatenate([got[:3 * i0], got[3 * i1:3 * N + i0], got[3 * N + i1:]])
        assert np.all(rest == -7.0), "rows of other ranks' particles must not be written"
        assert L.sfe_dev_peer_barrier(s) == 0                         # nobody rewrites a window that is still being read
        torch.cuda.synchronize()
    assert L.sfe_dev_bh_peer(win.data_ptr(), N, 24 * N, 28 * N, e.G, e.EPS2, e.TREE_THETA, e.WALK_EXACT, s) == e.STARFLOOD_FAILURE
    del win
    assert e.peer_close(dist.barrier) == 0, e.last_error()

    # ---- 2. Barnes-Hut leapfrog on `world` ranks against the oracle-driven run (bit for bit on the rank's slice)
    N, steps = 6000, 5
    pos, vel, mas = sfo.initcond(N)
    VX, VY = 160, 96
    e.peer_setup(rank, world, int(L.sfe_sim_peer_bytes_vis(N, 13, VX, VY)), allgather)
    sim = e.Sim(N); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas
    prm = e.sim_params(tree=True, nodes_max=2 * N)
    assert e.sim_upload(sim) == 0, e.last_error()
    sim.step_number = 0
    assert e.sim_solv_resident(sim, prm) == 0, e.last_error()
    rc, E = e.sim_step(sim, prm, nsteps=steps)
    assert rc == 0, e.last_error()
    # vis_draw() in multi-rank mode: every rank splats its slice, the integer buffers are summed over the
    # windows - every rank must end up with the frame the reference draws from the whole state, bit for bit, twice in a row
    p_, v_, a_, po_, *_rest = oracle_run(pos, vel, mas, steps, 2 * N)
    for frame in range(2):
        vis = e.Vis(); e.vis_init(vis, VX, VY)
        assert e.sim_vis_draw(vis, steps + frame) == 0, e.last_error()
        rbar, mvp, ab, rb = sfo.vis_draw(p_.copy(), po_.copy(), steps + frame, VX, VY)
        assert np.array_equal(vis.matMVP, mvp) and np.array_equal(vis.atomic_buffer, ab), (rank, "multi-rank vis_draw", frame)
    big = e.Vis(); e.vis_init(big, 4 * VX, 4 * VY)                      # a frame the window has no room for
    assert e.sim_vis_draw(big, steps) == e.STARFLOOD_FAILURE and "sfe_sim_peer_bytes_vis" in e.last_error()
    sim.mem[:] = -7.0
    assert e.sim_download(sim) == 0, e.last_error()
    p, v, a, po, ken, pen, Eo, box = oracle_run(pos, vel, mas, steps, 2 * N)
    i0, i1 = e.sim_slice(N, rank, world)
    assert np.array_equal(sim.pos, p) and np.array_equal(sim.mas, mas)          # every rank holds all positions
    for name, ref, w in (("vel", v, 3), ("acc", a, 3), ("pot", po, 1), ("ken", ken, 1), ("pen", pen, 1)):
        got = sim.find(name)
        assert np.array_equal(got[w * i0:w * i1], ref[w * i0:w * i1]), (rank, name)
        rest = np.concatenate([got[:w * i0], got[w * i1:]])
        assert np.all(rest == -7.0), (rank, name, "blocks of other ranks' particles must stay untouched")
    assert np.array_equal(e.sim_last_box(), box)
    assert np.all(np.abs(E - Eo) <= 1e-15 * np.abs(Eo).max()), (rank, E - Eo)
    allE = allgather(E.tobytes())
    assert all(x == allE[0] for x in allE), "every rank must report the same energies"
    # RAW snapshot: every rank writes the rows of its slice into ONE file; together they are the block
    import tempfile
    path = allgather(os.path.join(tempfile.gettempdir(), "sfe_peer_%d_%d.raw" % (port, os.getpid())))[0]
    assert e.sim_save_raw_async(path) == 0 and e.sim_save_wait() == 0, e.last_error()
    dist.barrier()
    want = np.concatenate([p, v, a, mas, po, ken, pen]).astype(np.float32)
    assert np.fromfile(path, np.float32).tobytes() == want.tobytes(), (rank, "snapshot file")
    dist.barrier()
    if rank == 0:
        os.remove(path)
    assert L.sfe_sim_release() == 0
    assert e.peer_close(dist.barrier) == 0, e.last_error()

    # ---- 3. PP + SPH on `world` ranks against the single-rank resident path.  A slice splits the source loop differently
    # from the full evaluation (same gate as tests/test_gpu_device_api.py: 1e-6 per solve), so this leg is held to the
    # parity tolerance of the path, 1e-5 rel-RMS, after three steps
    N, steps = 3000, 3
    pos, vel, mas, rad = sfo.initcond(N, outflow=0.0, sph=True)
    prm = e.sim_params(tree=False, sph=True)
    ref = e.Sim(N, sph=True); ref.pos[:] = pos; ref.vel[:] = vel; ref.mas[:] = mas; ref.rad[:] = rad
    assert e.sim_upload(ref) == 0 and e.sim_solv_resident(ref, prm) == 0, e.last_error()
    rc, Er = e.sim_step(ref, prm, nsteps=steps)
    assert rc == 0 and e.sim_download(ref) == 0, e.last_error()
    L.sfe_sim_release()
    e.peer_setup(rank, world, int(L.sfe_sim_peer_bytes(N, 16)), allgather)
    sim = e.Sim(N, sph=True); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas; sim.rad[:] = rad
    assert e.sim_upload(sim) == 0 and e.sim_solv_resident(sim, prm) == 0, e.last_error()
    rc, E = e.sim_step(sim, prm, nsteps=steps)
    assert rc == 0 and e.sim_download(sim) == 0, e.last_error()
    i0, i1 = e.sim_slice(N, rank, world)
    from conftest import rel_rms
    assert np.array_equal(sim.mas, ref.mas) and np.array_equal(sim.rad, ref.rad)
    for name in ("pos", "rho", "prs"):
        assert rel_rms(sim.find(name), ref.find(name)) <= 1e-5, (rank, name, rel_rms(sim.find(name), ref.find(name)))
    for name, w in (("vel", 3), ("acc", 3), ("pot", 1), ("ken", 1), ("pen", 1)):
        d = rel_rms(sim.find(name)[w * i0:w * i1], ref.find(name)[w * i0:w * i1])
        assert d <= 1e-5, (rank, name, d)
    assert np.all(np.abs(E - Er) <= 1e-6 * np.abs(Er).max()), (rank, E - Er)
    allE = allgather(E.tobytes()); allpos = allgather(sim.pos.tobytes())
    assert all(x == allE[0] for x in allE) and all(x == allpos[0] for x in allpos), "the ranks must agree bit for bit"
    # the same at a size where the SPH sums run over the cell list: every rank evaluates a dense range of the cell-sorted
    # targets, rho is stored into every rank's window and the pressure force into the owner's acc (routed over NVLink)
    assert L.sfe_sim_release() == 0 and e.peer_close(dist.barrier) == 0
    N, steps = 40000, 2
    pos, vel, mas, rad = sfo.initcond(N, outflow=0.0, sph=True)
    ref = e.Sim(N, sph=True); ref.pos[:] = pos; ref.vel[:] = vel; ref.mas[:] = mas; ref.rad[:] = rad
    assert e.sim_upload(ref) == 0 and e.sim_solv_resident(ref, prm) == 0, e.last_error()
    rc, Er = e.sim_step(ref, prm, nsteps=steps)
    assert rc == 0 and e.sim_download(ref) == 0, e.last_error()
    L.sfe_sim_release()
    e.peer_setup(rank, world, int(L.sfe_sim_peer_bytes(N, 16)), allgather)
    sim = e.Sim(N, sph=True); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas; sim.rad[:] = rad
    assert e.sim_upload(sim) == 0 and e.sim_solv_resident(sim, prm) == 0, e.last_error()
    rc, E = e.sim_step(sim, prm, nsteps=steps)
    assert rc == 0 and e.sim_download(sim) == 0, e.last_error()
    i0, i1 = e.sim_slice(N, rank, world)
    for name in ("pos", "rho", "prs"):
        assert rel_rms(sim.find(name), ref.find(name)) <= 1e-5, (rank, name, rel_rms(sim.find(name), ref.find(name)))
    for name, w in (("vel", 3), ("acc", 3), ("pot", 1)):
        d = rel_rms(sim.find(name)[w * i0:w * i1], ref.find(name)[w * i0:w * i1])
        assert d <= 1e-5, (rank, name, d)
    assert np.all(np.abs(E - Er) <= 1e-6 * np.abs(Er).max()), (rank, E - Er)
    allpos = allgather(sim.pos.tobytes()); allrho = allgather(sim.rho.tobytes())
    assert all(x == allpos[0] for x in allpos) and all(x == allrho[0] for x in allrho), "the ranks must agree bit for bit"
    # the window is too small for a bigger state: refused, and the engine falls back to nothing
    big = e.Sim(2 * N, sph=True); big.mas[:] = 1.0
    assert e.sim_upload(big) == e.STARFLOOD_FAILURE and "sfe_sim_peer_bytes" in e.last_error()
    assert e.peer_close(dist.barrier) == 0, e.last_error()
    dist.barrier()
    dist.destroy_process_group()
    e.shutdown()
    print("PEER_OK", rank, flush=True)


if __name__ == "__main__":
    main()
