"""Host-side logic that runs without a GPU: the numpy IC generator and root box against the oracle, the RAW
state layout, the target-slice partition, and the N > 1 data flow (all-gather of {x,y,z,m} slices) on the gloo
backend with world_size 2."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import sfo
from starflood_b200 import engine, ics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ics_match_reference_generator():
    for N in (1, 1000, 20000):
        p, v, m = ics.uniform_ball(N)
        P, V, M = sfo.initcond(N)
        assert np.array_equal(m, M)
        # identical up to libm rounding of cos/sin/acos/cbrt in numpy vs glibc (in practice: identical)
        assert np.abs(p - P).max() <= 1.2e-7 and np.abs(v - V).max() <= 1e-9
    p0, v0, m0 = ics.uniform_ball(4096, outflow=0.0)
    P0, V0, M0 = sfo.initcond(4096, outflow=0.0)
    assert np.abs(v0 - V0).max() <= 1e-9


def test_galaxy_pair_ics_match_the_oracle():
    for N in (1000, 50000):
        p, v, m = ics.galaxy_pair(N)
        P, V, M = sfo.initcond_galaxy_pair(N)
        assert np.array_equal(m, M) and np.abs(p - P).max() <= 2.4e-7 and np.abs(v - V).max() <= 1e-9


def test_pcg4d_matches_oracle():
    seeds = np.array([0x638FE17B, 0xB10BDCE7, 0x7B8ADB9B, 0xD8C5850E], np.uint32)
    want = sfo.pcg4d(seeds)
    got = ics.pcg4d(*[np.array([s], np.uint32) for s in seeds])
    assert [int(x[0]) for x in got] == [int(x) for x in want]
    assert [int(x) for x in engine._pcg4d(seeds)] == [int(x) for x in want]


@pytest.mark.parametrize("step", [0, 1, 999])
def test_root_box_matches_oracle(step):
    pos, vel, mas = sfo.initcond(30000)
    sim = engine.Sim(30000); sim.pos[:] = pos; sim.step_number = step
    assert np.array_equal(engine.root_box(sim), sfo.root_box(pos, step))


def test_sim_layout_is_raw_snapshot_layout():
    # blocks of N reals at offsets POS 0, VEL 3, ACC 6, MAS 9, POT 10, KEN 11, PEN 12, [RAD 13, RHO 14, PRS 15]
    N = 10
    sim = engine.Sim(N, sph=True)
    sim.mem[:] = np.arange(16 * N, dtype=np.float32)
    assert sim.pos[0] == 0 and sim.vel[0] == 3 * N and sim.acc[0] == 6 * N and sim.mas[0] == 9 * N
    assert sim.pot[0] == 10 * N and sim.rad[0] == 13 * N and sim.rho[0] == 14 * N and sim.prs[0] == 15 * N
    assert engine.Sim(N).mem.size == 13 * N
    r = sfo.RefSim("pp", 7) if sfo.ref_available("pp") else None
    if r is not None:
        assert r.mem.size == 13 * 7


def test_slices_partition_targets():
    for N in (1, 10, 1 << 20, 12345):
        for G in (1, 2, 4, 8):
            cuts = [(N * g) // G for g in range(G + 1)]
            assert cuts[0] == 0 and cuts[-1] == N and all(b >= a for a, b in zip(cuts, cuts[1:]))


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from oracle import sfo
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N = 3000
pos, vel, mas = sfo.initcond(N)
pm = np.concatenate([pos.reshape(-1, 3), mas[:, None]], axis=1).astype(np.float32)
i0, i1 = (N * rank) // world, (N * (rank + 1)) // world
mine = torch.from_numpy(pm[i0:i1].copy())
# ragged slices: gather with explicit sizes (N is not a multiple of world in general)
sizes = [((N * (g + 1)) // world - (N * g) // world) for g in range(world)]
parts = [torch.empty(s, 4) for s in sizes]
dist.all_gather(parts, mine)
full = torch.cat(parts).numpy()
assert np.array_equal(full, pm)
# each rank evaluates its own target slice over all sources with the CPU restatement; together they tile the answer
acc, pot = sfo.pp(np.ascontiguousarray(full[:, :3]).ravel(), np.ascontiguousarray(full[:, 3]), i0, i1)
out = torch.from_numpy(np.concatenate([acc.reshape(-1, 3)[i0:i1], pot[i0:i1, None]], axis=1).copy())
outs = [torch.empty(s, 4) for s in sizes]
dist.all_gather(outs, out)
got = torch.cat(outs).numpy()
ref_acc, ref_pot = sfo.pp(pos, mas)
assert np.array_equal(got[:, :3].ravel(), ref_acc) and np.array_equal(got[:, 3], ref_pot)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("N,world", [(1 << 20, 8), (1000, 3), (5, 8), (0, 4), (6000, 16), (30011, 7)])
def test_sim_slice_is_a_partition_and_names_the_owner(N, world):
    """rank r owns [r*N/world, (r+1)*N/world).  The routed walk (csrc/sfe_walk.cu) finds the owner of a
    particle from its index with the same arithmetic; mirrored here."""
    from starflood_b200 import engine
    cuts = [engine.sim_slice(N, r, world) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == N
    assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
    assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
    for i in list(range(min(N, 50))) + list(range(max(N - 50, 0), N)) + ([N // 2, N // 3] if N else []):
        r = (i * world) // N
        while i < (N * r) // world:
            r -= 1
        while i >= (N * (r + 1)) // world:
            r += 1
        assert cuts[r][0] <= i < cuts[r][1]


def test_two_rank_gloo_slice_flow(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_device_libm_restatement_matches_host_libm():
    """csrc/sfe_libm.cuh (the logf / sinf / cosf the splat kernel evaluates, glibc's algorithms in IEEE double operations)
    against the host C library: tools/libm_probe.cpp compares bit patterns over every `stride`-th float of the ranges
    vis_draw() can produce (stride 1 = exhaustive, 4.4e9 values, 0 mismatches, run by hand; a prime stride here)."""
    exe = os.path.join(ROOT, "tools", "libm_probe")
    src = os.path.join(ROOT, "tools", "libm_probe.cpp")
    hdr = os.path.join(ROOT, "starflood_b200", "csrc", "sfe_libm.cuh")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fopenmp", src, "-o", exe])
    import json
    out = json.loads(subprocess.check_output([exe, "211"]).decode())
    for f in ("logf", "cosf", "sinf"):
        assert out[f]["checked"] > 10_000_000 and out[f]["mismatches"] == 0, out
        assert out[f]["correctly_rounded_differs"] > 0      # why a correctly rounded routine would not do


@pytest.mark.parametrize("step", [0, 5, 4000])
def test_vis_model_view_matches_oracle(step):
    """engine.vis_model_view() (host side of vis_draw(), visualization.c:344-446) equals the oracle's matrices."""
    vm, mvp = sfo.vis_matrices(0.75, step)
    assert np.array_equal(engine.vis_model_view(step), vm)
