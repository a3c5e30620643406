"""Multi-rank path over the engine's own peer exchange (sfe_peer_*, csrc/sfe_peer.cu): one process per
rank, CUDA IPC windows, drift + all-gather fused into one kernel, rho exchange between the SPH passes, {T, V} combined
in rank order.  The ranks run in tests/peer_worker.py; on a single-GPU box they share device 0."""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_world(world, timeout=150, **extra_env):
    port = free_port()
    env = dict(os.environ, SFE_PEER_TIMEOUT_MS="60000", **extra_env)
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "peer_worker.py"), str(r), str(world), str(port)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=timeout)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "PEER_OK %d" % r in o, "rank %d failed:\n%s" % (r, o[-3000:])


def test_two_ranks_exchange_and_leapfrog():
    run_world(2)


def test_three_ranks_uneven_slices():
    run_world(3)


@pytest.mark.parametrize("kernel", ["0", "2"], ids=["walk_kernel", "walk_p_kernel"])
def test_two_ranks_with_the_throughput_walk_kernels(kernel):
    """The ranks' shares are small here, so the routed walk normally runs walk_sp_kernel; SFE_WALK_SP forces the two
    throughput kernels (what a rank runs at large N) through the same bit-for-bit leapfrog checks."""
    run_world(2, SFE_WALK_SP=kernel)


def test_single_rank_window_and_errors(engine):
    import torch
    L = engine.lib()
    s = torch.cuda.current_stream().cuda_stream
    h = C.create_string_buffer(64)
    assert L.sfe_peer_close() == 0
    assert L.sfe_peer_open(h) == engine.STARFLOOD_FAILURE                       # no window yet
    assert L.sfe_dev_peer_barrier(s) == engine.STARFLOOD_FAILURE
    assert L.sfe_peer_create(2, 2, 1024, h) == engine.STARFLOOD_FAILURE         # rank outside the world
    assert L.sfe_peer_create(0, 17, 1024, h) == engine.STARFLOOD_FAILURE        # more than SFE_PEER_MAX ranks
    assert L.sfe_peer_create(0, 1, 1000, h) == 0, engine.last_error()
    assert h.raw != bytes(64) and L.sfe_peer_window()
    assert L.sfe_dev_peer_put(0, 16, s) == engine.STARFLOOD_FAILURE             # created but not open
    assert L.sfe_peer_open(h) == 0, engine.last_error()
    assert L.sfe_peer_open(h) == engine.STARFLOOD_FAILURE                       # already open
    for k in range(5):
        assert L.sfe_dev_peer_put(8 * k, 100, s) == 0, engine.last_error()      # a world of one: flags only
    assert L.sfe_dev_peer_put(1000, 100, s) == engine.STARFLOOD_FAILURE         # outside the (256-rounded) window
    torch.cuda.synchronize()
    # a one-rank window does not switch the resident leapfrog into multi-rank mode
    N = 500
    from oracle import sfo
    pos, vel, mas = sfo.initcond(N)
    sim = engine.Sim(N); sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas
    prm = engine.sim_params(tree=True, nodes_max=2 * N)
    assert engine.sim_upload(sim) == 0 and engine.sim_solv_resident(sim, prm) == 0
    rc, E = engine.sim_step(sim, prm, nsteps=2)
    assert rc == 0 and engine.sim_download(sim) == 0
    assert np.isfinite(sim.mem).all() and not np.array_equal(sim.pos, pos)
    assert L.sfe_sim_release() == 0 and L.sfe_peer_close() == 0
    assert L.sfe_sim_peer_bytes(1000, 13) == 48 * 1024 and L.sfe_sim_peer_bytes(1000, 16) == 52 * 1024
