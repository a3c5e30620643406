#!/usr/bin/env python
"""bench.py - headline benchmark of the force-evaluation hot path (BASELINE.json).

Workload (config.workload): C2 = Barnes-Hut gravity, N = 2^20 uniform sphere from the reference's
initcond_generate(), theta = 0.5 monopole, root box of sim_solv().  One "step" = one pass of the hot path over
one snapshot: the SOLVER_TREE branch of sim_solv() (simulation.c:245-325) - root box of the step, tree_build,
solve_gravity_part_tree for all N particles - which is exactly what the reference arm times.

    value   particle-steps/s with the snapshot already resident in HBM (device entry points, sfe_dev_*)
    e2e     the same through the host C ABI (sfe_solv_tree on pinned HOST buffers): every step copies
            pos/mas host->device and acc/pot device->host inside the timed region
    pp      PP direct-sum interactions/s at the same N (second half of BASELINE.json's metric), FP32-bound
    roofline  the walk kernel against the roofs that bind it (instruction issue, L1 data pipe), with SURVEY 8(d)'s HBM byte model beside it
    cpu_baseline  the unmodified reference (oracle/_ref/libsfref_tree.so) timed on this box's host cores

N > 1 (torchrun, one rank per GPU): strong scaling at the same total N.  Every rank owns a target slice,
all-gathers the {x,y,z,m} records over NCCL/NVLink (torch.distributed), builds the full tree locally and walks
its own slice.  Time is the max over ranks.

Timing hygiene: W >= 3 warm-up steps; the timed steps rotate through a ring of distinct snapshots whose total
footprint exceeds L2 (config.l2), so no step finds its inputs cached; CUDA events on the launching stream;
nvidia-smi clocks sampled during the timed region.

--impl reference times the reference's own CPU implementation (all host threads) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DEFAULT = 1 << 20
THETA = 0.5
L2_BYTES = 126 << 20
# distinct node visits per particle of the reference walk on C2 (opened + evaluated + empty),
# re-measured by the oracle in tests/test_gpu_walk.py::test_full_size
V_C2 = 514.3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--flavour", default="exact", choices=["exact", "fast"])
    ap.add_argument("--no-pp", action="store_true", help="skip the PP leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sph", action="store_true", help="skip the C3 (PP + SPH) legs")
    ap.add_argument("--no-resident", action="store_true", help="skip the device-resident leapfrog legs")
    ap.add_argument("--no-vis", action="store_true", help="skip the vis_draw() leg")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 leg (PP N=2^22 by target slice; runs on 2/4/8 GPUs)")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 leg (Barnes-Hut N=2^24 galaxy pair)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: all-gather of {x,y,z,m} by the engine's peer-memory kernels (sfe_dev_peer_put) or by NCCL")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ helpers
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first(self, timeout=8.0):
        """Blocks until nvidia-smi has delivered its first row (its start-up takes 0.1 - 2 s on a fresh box)."""
        t_end = time.time() + timeout
        while self.proc and not self.rows and time.time() < t_end:
            time.sleep(0.02)

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if not (self.t0 - 0.05 <= ts <= self.t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": max(pw),
                "reasons": sorted(reasons), "samples": len(sm)}


def snapshots(N, count):
    """Ring of distinct snapshots: the reference ICs rotated about z by different angles (same density profile and
    walk statistics), each with the root box sim_solv() would compute at step_number = k."""
    from starflood_b200 import engine as e, ics
    pos, vel, mas = ics.uniform_ball(N)
    p = pos.reshape(-1, 3)
    out = []
    for k in range(count):
        a = 0.37 * k
        c, s = np.float32(np.cos(a)), np.float32(np.sin(a))
        q = np.empty_like(p)
        q[:, 0] = c * p[:, 0] - s * p[:, 1]; q[:, 1] = s * p[:, 0] + c * p[:, 1]; q[:, 2] = p[:, 2]
        sim = e.Sim(N, mem=None)
        sim.pos[:] = q.ravel(); sim.mas[:] = mas; sim.step_number = k
        out.append((sim, e.root_box(sim)))
    return out


def ref_lib_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsfref_tree.so"))


def workload_name(N):
    """config.workload, the same string on the engine arm and on the reference arm (the driver compares them)."""
    return ("C2: Barnes-Hut gravity N=%d uniform sphere (reference initcond), theta=0.5 monopole, "
            "one sim_solv() per step: root box + tree_build + solve_gravity_part_tree" % N)


def cpu_reference_steps(N, steps, warmup):
    """The reference's own sim_solv() (root box + tree_build + solve_gravity_part_tree) from oracle/_ref on all
    host threads; falls back to the oracle port when the reference library did not travel.
    Returns (kind, threads actually used, host cores, per-step seconds)."""
    from oracle import sfo
    cores = sfo.host_cores()
    # README.md:242-248 of the reference: OMP_NUM_THREADS=$(nproc) OMP_PROC_BIND=true.  Set unconditionally - a launcher's
    # own OMP_NUM_THREADS (torch.distributed.run exports 1) must not win - and, because libgomp may already be loaded
    # and reads the environment only once, also through the runtime's API; the count reported is the one read back.
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ["OMP_PROC_BIND"] = "true"
    times = []
    if ref_lib_available() and N <= 2 * sfo.REF_NODES_MAX["tree"]:
        kind = "reference"
        r = sfo.RefSim("tree", N)
        threads = sfo.omp_threads(cores)
        r.initcond()
        for s in range(warmup + steps):
            t0 = time.perf_counter(); r.solv(s); dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
        r.close()
    else:
        kind = "port"
        threads = 1                       # the oracle port is scalar C
        pos, vel, mas = sfo.initcond(N)
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            b = sfo.root_box(pos, s)
            rc, nodes, n = sfo.tree_build_insert(b, pos, mas, N)
            sfo.bh_walk(nodes[:n], pos, mas)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
    return kind, threads, cores, times


def host_cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return None


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import sfo
    N = args.n
    kind, threads, cores, times = cpu_reference_steps(N, args.steps, args.warmup)
    ms = 1e3 * float(np.mean(times))
    value = N / (ms * 1e-3)
    sample = ("%d full sim_solv() steps (root box + tree_build + BH walk) at N=%d after %d warm-up, %d OpenMP threads "
              "(omp_get_max_threads) on %d host cores" % (args.steps, N, args.warmup, threads, cores))
    line = {
        "impl": "reference", "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(N), "N": N, "theta": THETA, "depth_max": 16,
                   "arm": "the unmodified reference sim_solv() on the host CPUs (OpenMP), the same IC re-solved with the "
                          "root box of step_number = 0, 1, 2, ... (the engine arm rotates z-rotated copies of the same IC: "
                          "statistically the same work)"},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": threads, "host_cores": cores,
                         "cpu_model": host_cpu_model(), "kind": kind, "sample": sample,
                         "build_flags": sfo.REF_BUILD_FLAGS + " (BASELINE.md section 3 asks for -march=native; the library is "
                                        "built in the dev container for a different host CPU, hence x86-64-v3)"},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ secondary legs
C3_N = 262144


def sph_leg(e, L, torch, dev, stream):
    """C3 snapshot: N = 262144 gas sphere, h = 0.1.  Times density + EOS + pressure force (the solve_sph()
    replacement) with the snapshot resident, and reports it against SURVEY 8(d)'s byte model with the cutoff the cell
    list uses (R = 4 h): B = 20 B n_rho + 24 B n_F + 64 B per particle, n = mean number of sources within R."""
    from starflood_b200 import ics
    N = C3_N
    pos, vel, mas = ics.uniform_ball(N, outflow=0.0)
    rad = np.full(N, 0.1, np.float32)
    pm = np.empty((N, 4), np.float32); pm[:, :3] = pos.reshape(-1, 3); pm[:, 3] = mas
    d_pm4 = torch.from_numpy(pm).to(dev).reshape(-1)
    d_rad = torch.from_numpy(rad).to(dev)
    rho = torch.zeros(N, dtype=torch.float32, device=dev); prs = torch.zeros_like(rho)
    acc = torch.zeros(3 * N, dtype=torch.float32, device=dev)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
    sptr = stream.cuda_stream

    def solve():
        rc = L.sfe_dev_sph_density(d_pm4.data_ptr(), d_rad.data_ptr(), N, 0, N, rho.data_ptr(), sptr)
        rc |= L.sfe_dev_sph_eos(rho.data_ptr(), prs.data_ptr(), N, e.EOS_K, e.EOS_N, sptr)
        rc |= L.sfe_dev_sph_force(d_pm4.data_ptr(), d_rad.data_ptr(), rho.data_ptr(), prs.data_ptr(), N, 0, N, e.SPH_CUTOFF2,
                                  acc.data_ptr(), sptr)
        if rc != 0:
            raise SystemExit("bench.py: SPH call failed: " + e.last_error())
    import ctypes as C
    pairs = C.c_uint64()
    for _ in range(3):
        solve()
    torch.cuda.synchronize()
    L.sfe_sph_pairs_evaluated(C.byref(pairs), sptr)     # reset the counter
    solve()
    L.sfe_sph_pairs_evaluated(C.byref(pairs), sptr)     # (target, source) pairs the two cell-list passes of ONE solve evaluated
    pairs_eval = float(pairs.value)
    reps, tot = 10, 0.0
    for _ in range(reps):
        flush.fill_(1)                                  # 160 MiB > L2: the snapshot is read from HBM again
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); solve(); b.record(stream)
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    ms = tot / reps
    # mean number of sources within R = 4 h of a target, on a sample (input analysis, outside every timed region)
    from scipy.spatial import cKDTree
    P = pos.reshape(-1, 3).astype(np.float64)
    sample = P[np.random.default_rng(1).choice(N, 2048, replace=False)]
    nn = float(np.mean(cKDTree(P).query_ball_point(sample, 0.4, return_length=True)))
    hbm, hbm_src = measured_peaks()
    bytes_alg = (20.0 * nn + 24.0 * nn + 64.0) * N
    ach = bytes_alg / (ms * 1e-3) / 1e9
    # the bound the pair passes actually run into: FP32 FMA-pipe lane operations and MUFU.EX2 per EVALUATED pair (one common h:
    # density 3 add + 2 mul + 4 fma = 9, force 4 add + 4 mul + 5 fma + 2 select = 15; one ex2 each), against
    # 128 FMA lanes and 16 MUFU lanes per SM per clock
    info = e.device_info()
    clk = info["clock_khz"] * 1e3
    fma_peak, mufu_peak = 128.0 * info["sm_count"] * clk, 16.0 * info["sm_count"] * clk
    pairs_in = 2.0 * nn * N
    compute = {"bound": "fp32 fma pipe", "pairs_evaluated_per_solve": pairs_eval, "pairs_in_range_per_solve": pairs_in,
               "evaluated_over_in_range": pairs_eval / pairs_in if pairs_in else None,
               "pairs_per_s": pairs_eval / (ms * 1e-3),
               "fma_lane_ops_per_pair": 12.0, "fma_frac": 12.0 * pairs_eval / (ms * 1e-3) / fma_peak,
               "mufu_per_pair": 1.0, "mufu_frac": pairs_eval / (ms * 1e-3) / mufu_peak,
               "useful_fma_frac": 12.0 * pairs_in / (ms * 1e-3) / fma_peak,
               "note": "evaluated pairs are counted on the device (sfe_sph_pairs_evaluated); in-range = sources within R = 4 h of a target "
                       "(sampled); the whole solve_sph time (grid, sort, records, two pair passes, EOS) is charged to the pair work"}
    return {"workload": "C3 snapshot: SPH density + EOS + pressure force, N=%d, h=0.1 (solve_sph replacement), device resident" % N,
            "ms_per_solve": ms, "particle_solves_per_s": N / (ms * 1e-3),
            "compute_roofline": compute,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "peak_source": hbm_src,
                         "model": "SURVEY 8(d): (20 B + 24 B) x n + 64 B per particle, n = %.0f sources within R = 4 h = 0.4 "
                                  "(sampled); source tiles are shared through shared memory, so DRAM traffic is far below this" % nn,
                         "algorithmic_bytes_per_solve": bytes_alg},
            "l2": "160 MiB flush buffer written between repetitions"}


def vis_leg(e, L, torch, dev, stream, N, pm4, pot, with_cpu):
    """vis_draw() replacement (visualization.c:253-739) on the C2 state: N particles x SPATIAL_SAMPLES 128
    samples into the reference's default 1080 x 1080 image (main.c:81).  `ms_per_frame` = mean radius + matrix + clear +
    splat + finishing pass with the state resident in HBM; `e2e` = sfe_vis_draw() on pinned host arrays (pos + pot up,
    render buffer down); `cpu_reference` = the unmodified reference's vis_draw() on the host cores."""
    import ctypes as C
    sx = sy = 1080
    npix = sx * sy
    vm = e.vis_model_view(0)
    ortho, auto, samples, scale, tonemap = e._vis_args()
    ab = torch.zeros(4 * npix, dtype=torch.int32, device=dev)
    rb = torch.zeros(4 * npix, dtype=torch.float32, device=dev)
    info = torch.zeros(32, dtype=torch.float64, device=dev)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
    sptr = stream.cuda_stream

    def frame():
        if L.sfe_dev_vis_draw(pm4.data_ptr(), pot.data_ptr(), N, 0, vm.ctypes.data, ortho, auto, sx, sy, samples, scale, tonemap,
                              ab.data_ptr(), rb.data_ptr(), info.data_ptr(), sptr) != 0:
            raise SystemExit("bench.py: sfe_dev_vis_draw failed: " + e.last_error())
    for _ in range(3):
        frame()
    torch.cuda.synchronize()
    reps, tot = 10, 0.0
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); frame(); b.record(stream)
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    ms = tot / reps
    landed = int(ab[3::4].sum().item())
    # through the host entry point (what the vis shim calls), pinned host arrays
    host_pm = pm4.reshape(-1, 4).cpu().numpy()
    sim = e.Sim(N, pinned=True)
    sim.pos[:] = host_pm[:, :3].reshape(-1); sim.mas[:] = host_pm[:, 3]; sim.pot[:] = pot.cpu().numpy()
    vis = e.Vis(); e.vis_init(vis, sx, sy)
    vis.atomic_buffer = None                              # like the shim: the accumulation buffer stays on the device
    e._check(L.sfe_host_register(vis.render_buffer.ctypes.data, vis.render_buffer.nbytes), "sfe_host_register")
    for _ in range(2):
        e._check(e.vis_draw(vis, sim), "sfe_vis_draw")
    t0 = time.perf_counter()
    for _ in range(5):
        e._check(e.vis_draw(vis, sim), "sfe_vis_draw")
    e2e_ms = (time.perf_counter() - t0) / 5 * 1e3
    L.sfe_host_unregister(vis.render_buffer.ctypes.data)
    out = {"workload": "vis_draw(): N=%d particles x %d samples into %d x %d (reference defaults), on the C2 state" % (N, samples, sx, sy),
           "ms_per_frame": ms, "samples_per_s": float(N) * samples / (ms * 1e-3), "samples_landed": landed,
           "e2e": {"ms_per_frame": e2e_ms, "h2d_bytes": 16 * N, "d2h_bytes": 16 * npix, "api": "sfe_vis_draw on pinned host arrays"},
           "parity": "integer accumulation buffer bit-identical to the reference (tests/test_gpu_vis.py)",
           "l2": "160 MiB flush buffer written between repetitions"}
    if with_cpu:
        from oracle import sfo
        if sfo.ref_available("vis"):
            r = sfo.RefSim("vis", N)
            r.pos[:] = sim.pos; r.mas[:] = sim.mas; r.pot[:] = sim.pot
            cores = sfo.omp_threads(sfo.host_cores())
            r.vis_draw(0, sx, sy)
            t = []
            for _ in range(2):
                t0 = time.perf_counter(); r.vis_draw(0, sx, sy); t.append(time.perf_counter() - t0)
            r.close()
            out["cpu_reference"] = {"ms_per_frame": 1e3 * min(t), "cores": cores, "kind": "reference",
                                    "sample": "2 full vis_init()+vis_draw() calls of the unmodified visualization.c at the same N "
                                              "and image size (best), OpenMP"}
    return out


def resident_leg(e, L, torch, dist, world, rank, N, tree, with_sph, flavour, allgather_bytes, barrier, max_over_ranks, use_peer):
    """Whole sim_step()s (kick, drift, root box, build + walk | PP, [SPH], kick, energies) on the HBM-resident state
    through sfe_sim_run(); with more than one rank the drift kernel stores straight into every rank's window."""
    from starflood_b200 import ics
    pos, vel, mas = ics.uniform_ball(N, outflow=4.0e-3 if tree else 0.0)
    sim = e.Sim(N, sph=with_sph)
    sim.pos[:] = pos; sim.vel[:] = vel; sim.mas[:] = mas
    if with_sph:
        sim.rad[:] = 0.1
    if world > 1:
        if not use_peer:
            return None                              # the resident leapfrog's exchange is the engine's own
        e.peer_setup(rank, world, int(L.sfe_sim_peer_bytes(N, sim.dof)), allgather_bytes)
    prm = e.sim_params(tree=tree, sph=with_sph, flavour=flavour, nodes_max=max(N, 1024))
    if e.sim_upload(sim) != 0 or e.sim_solv_resident(sim, prm) != 0:
        raise SystemExit("bench.py: resident leg failed: " + e.last_error())
    steps = 20 if tree else 5
    sim.step_number = 0
    rc, _ = e.sim_step(sim, prm, nsteps=3)
    barrier()
    l0 = L.sfe_launch_count()
    t0 = time.perf_counter()
    sim.step_number = 3
    rc2, E = e.sim_step(sim, prm, nsteps=steps)
    barrier()
    ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / steps)
    if rc != 0 or rc2 != 0:
        raise SystemExit("bench.py: resident leg failed: " + e.last_error())
    launches = int(L.sfe_launch_count() - l0)
    L.sfe_sim_release()
    if world > 1:
        e.peer_close(dist.barrier)
    return {"workload": ("Barnes-Hut" if tree else "PP + SPH") + " leapfrog N=%d, %d sim_step()s through sfe_sim_run()" % (N, steps),
            "ms_per_step": ms, "particle_steps_per_s": N / (ms * 1e-3), "gpu_launches": launches,
            "energy_last": [float(x) for x in E[-1]],
            "timing": "host clock around the synchronous call, barrier on both sides, max over ranks",
            "l2": "the state evolves in place (a real trajectory); its working set is what the workload has"}


def slot_hash(torch, acc, pot):
    """Order-independent 64-bit checksum of the bit patterns of acc / pot (the EXACT walk is bit-reproducible, so the
    value is the same at every GPU count when every rank's slice is right)."""
    a = acc.view(torch.int32).to(torch.int64).reshape(-1, 3)
    p = pot.view(torch.int32).to(torch.int64)
    h = (a[:, 0] * 1000003 + a[:, 1]) * 1000003 + a[:, 2]
    h = h * 1000003 + p
    return int((h ^ (h >> 29)).sum().item()) & 0xFFFFFFFFFFFFFFFF


def c5_leg(e, L, torch, dist, world, rank, dev, stream, flavour, barrier, max_over_ranks):
    """BASELINE.json C5: Barnes-Hut N = 2^24 galaxy-pair collision ICs, theta = 0.5, tree per GPU.  A step = root box +
    full tree build on every rank + walk of the rank's target slice (index slices: sfe_dev_bh).  Parity: rank 0 walks ALL
    particles on its own tree afterwards and compares the bits of every rank's slice with it."""
    from starflood_b200 import ics
    N = 1 << 24
    nodes_max = 1 << 23
    sptr = stream.cuda_stream
    pos, vel, mas = ics.galaxy_pair(N)
    pm = np.empty((N, 4), np.float32); pm[:, :3] = pos.reshape(-1, 3); pm[:, 3] = mas
    del pos, vel
    pm4 = torch.from_numpy(pm).to(dev).reshape(-1)
    del pm
    i0, i1 = (N * rank) // world, (N * (rank + 1)) // world
    ni = i1 - i0
    acc = torch.zeros(3 * N, dtype=torch.float32, device=dev)
    pot = torch.zeros(N, dtype=torch.float32, device=dev)
    box = np.zeros(6, np.float32)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def step(k, timed=False, lo=i0, n=ni):
        if timed: ev[0].record(stream)
        rc = L.sfe_dev_tree_build_box(pm4.data_ptr(), N, k, e.TREE_DEPTH_MAX, nodes_max, sptr)
        if timed: ev[1].record(stream)
        rc |= L.sfe_dev_bh(pm4.data_ptr(), N, lo, n, acc.data_ptr(), pot.data_ptr(), e.G, e.EPS2, THETA, flavour, sptr)
        if timed: ev[2].record(stream)
        if rc != 0:
            raise SystemExit("bench.py: C5 leg failed: " + e.last_error())
    step(0); step(1)
    barrier()
    import ctypes as C
    st = (C.c_uint32(), C.c_uint32(), C.c_uint32())
    if L.sfe_dev_tree_status(sptr, C.byref(st[0]), C.byref(st[1]), C.byref(st[2])) != 0:
        raise SystemExit("bench.py: C5 tree build failed: " + e.last_error())
    reps = 3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    build_ms = walk_ms = 0.0
    l0 = L.sfe_launch_count()
    a.record(stream)
    for r in range(reps):
        step(2 + r, timed=True)
        stream.synchronize()
        build_ms += ev[0].elapsed_time(ev[1]); walk_ms += ev[1].elapsed_time(ev[2])
    b.record(stream)
    barrier()
    launches = int(L.sfe_launch_count() - l0)
    ms = max_over_ranks(a.elapsed_time(b)) / reps
    build_ms = max_over_ranks(build_ms / reps); walk_ms = max_over_ranks(walk_ms / reps)
    # parity on the last snapshot's box: every rank's slice against a single-rank walk of all particles on rank 0
    mine_acc = acc[3 * i0:3 * i1].clone(); mine_pot = pot[i0:i1].clone()
    parity = None
    if world > 1:
        even = N % world == 0
        ga = [torch.empty_like(mine_acc) for _ in range(world)] if even else None
        gp = [torch.empty_like(mine_pot) for _ in range(world)] if even else None
        if even:
            dist.all_gather(ga, mine_acc); dist.all_gather(gp, mine_pot)
        if rank == 0 and even:
            step(2 + reps - 1, lo=0, n=N)
            torch.cuda.synchronize()
            ok = bool(torch.equal(torch.cat(ga).view(torch.int32), acc.view(torch.int32)) and
                      torch.equal(torch.cat(gp).view(torch.int32), pot.view(torch.int32)))
            parity = {"bit_identical_to_single_rank": ok, "checked": "acc and pot of all %d particles, gathered from the %d ranks, "
                      "against a walk of all particles on rank 0's own tree" % (N, world),
                      "checksum": "%016x" % slot_hash(torch, acc, pot)}
    else:
        parity = {"bit_identical_to_single_rank": True, "checked": "single rank (this value is the one the multi-GPU runs must reproduce)",
                  "checksum": "%016x" % slot_hash(torch, acc, pot)}
    hbm, hbm_src = measured_peaks()
    out = {"workload": "C5: Barnes-Hut gravity N=%d galaxy-pair collision ICs (initcond.c:265-329), theta=0.5 monopole, tree per GPU; "
                       "one sim_solv() per step: root box + tree_build + solve_gravity_part_tree" % N,
           "N": N, "ms_per_step": ms, "particle_steps_per_s": N / (ms * 1e-3), "build_ms": build_ms, "walk_ms": walk_ms,
           "tree_nodes": int(st[0].value), "tree_depth": int(st[2].value), "gpu_launches": launches, "steps": reps,
           "partition": "target index slices, full tree on every GPU" if world > 1 else "single GPU",
           "parity": parity,
           "build_hbm": {"algorithmic_bytes": (16.0 + 24.0 * int(st[2].value)) * N + 32.0 * int(st[0].value),
                         "achieved_gbs": ((16.0 + 24.0 * int(st[2].value)) * N + 32.0 * int(st[0].value)) / (build_ms * 1e-3) / 1e9,
                         "frac": ((16.0 + 24.0 * int(st[2].value)) * N + 32.0 * int(st[0].value)) / (build_ms * 1e-3) / 1e9 / hbm,
                         "model": "SURVEY 8(d): B_build = 16 + 24 D + 32 nodes/N bytes per particle", "peak_source": hbm_src},
           "l2": "inputs of 256 MiB > L2; every step rebuilds the tree (a new root box per step)"}
    del pm4, acc, pot
    torch.cuda.empty_cache()
    return out


def c4_leg(e, L, torch, dist, world, rank, dev, stream, barrier, max_over_ranks):
    """BASELINE.json C4: PP direct-sum gravity N = 2^22 sharded by target slice.  Every rank evaluates its slice over all
    sources (sfe_dev_pp).  Parity: rank 0 re-evaluates the first 256 targets of every rank's slice on its own and compares
    (the j-split of the kernel depends on the number of targets, so the comparison is to 1e-6 relative, not bitwise)."""
    from starflood_b200 import ics
    N = 1 << 22
    sptr = stream.cuda_stream
    pos, vel, mas = ics.uniform_ball(N)
    pm = np.empty((N, 4), np.float32); pm[:, :3] = pos.reshape(-1, 3); pm[:, 3] = mas
    pm4 = torch.from_numpy(pm).to(dev).reshape(-1)
    i0, i1 = (N * rank) // world, (N * (rank + 1)) // world
    ni = i1 - i0
    acc = torch.zeros(3 * N, dtype=torch.float32, device=dev)
    pot = torch.zeros(N, dtype=torch.float32, device=dev)

    def solve(lo, n):
        if L.sfe_dev_pp(pm4.data_ptr(), N, lo, n, acc.data_ptr(), pot.data_ptr(), e.G, e.EPS2, sptr) != 0:
            raise SystemExit("bench.py: C4 leg failed: " + e.last_error())
    solve(i0, min(ni, 65536))                                  # warm-up on a small slice (the full one takes a second)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); solve(i0, ni); b.record(stream)
    barrier()
    ms = max_over_ranks(a.elapsed_time(b))
    inter = float(N) * float(N) / (ms * 1e-3)
    info = e.device_info()
    K = 256
    head_a = acc[3 * i0:3 * (i0 + K)].clone(); head_p = pot[i0:i0 + K].clone()
    ga = [torch.empty_like(head_a) for _ in range(world)]; gp = [torch.empty_like(head_p) for _ in range(world)]
    if world > 1:
        dist.all_gather(ga, head_a); dist.all_gather(gp, head_p)
    else:
        ga, gp = [head_a], [head_p]
    parity = None
    if rank == 0:
        worst = 0.0
        for r in range(world):
            lo = (N * r) // world
            solve(lo, K)
            torch.cuda.synchronize()
            ra = acc[3 * lo:3 * (lo + K)].double(); rp = pot[lo:lo + K].double()
            worst = max(worst, float(((ga[r].double() - ra).norm() / ra.norm()).item()), float(((gp[r].double() - rp).norm() / rp.norm()).item()))
        parity = {"max_rel_rms_vs_rank0": worst, "tolerance": 1e-6, "ok": worst <= 1e-6,
                  "checked": "the first %d targets of every rank's slice, re-evaluated on rank 0" % K}
    out = {"workload": "C4: PP direct-sum gravity N=%d sharded by target slice, one solve_gravity_part_part per step" % N,
           "N": N, "ms_per_solve": ms, "interactions_per_s": inter, "flop_per_interaction": 20,
           "achieved_tflops": 20.0 * inter / 1e12, "fp32_peak_tflops": info["fp32_peak_tflops"] * world,
           "frac": 20.0 * inter / 1e12 / (info["fp32_peak_tflops"] * world), "parity": parity,
           "partition": "target index slices, all sources on every GPU", "steps": 1}
    del pm4, acc, pot
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------ engine arm
def run_engine(args):
    import torch
    import torch.distributed as dist
    from starflood_b200 import engine as e

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    e.init(local)
    L = e.lib()
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    flavour = e.WALK_EXACT if args.flavour == "exact" else e.WALK_FAST

    N = args.n
    G = world
    # target slices: rank g owns [g*N/G, (g+1)*N/G)
    i0 = (N * rank) // G
    i1 = (N * (rank + 1)) // G
    ni = i1 - i0
    ring = max(3, (2 * L2_BYTES) // (16 * N) + 1)
    snaps = snapshots(N, ring)
    nodes_max = max(N, 1024)

    def allgather_bytes(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    class Window:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False), "version": 2}

    # device-resident state: per snapshot the rank's slice of pm4 records; full pm4 + outputs shared
    use_peer = world > 1 and args.exchange == "peer"
    peer_note = None
    if use_peer:
        # two {x,y,z,m} arrays inside the exchange window, alternated between steps (sfengine.h: a range written by
        # exchange k may be rewritten by exchange k+2 at the earliest)
        # ... and acc / pot: the routed walk (sfe_dev_bh_peer) stores every result into its owner's window
        ok = 1
        try:
            wptr = e.peer_setup(rank, world, 3 * 16 * N, allgather_bytes)
        except e.EngineError as err:      # e.g. CUDA IPC not permitted in this container: every rank falls back together
            ok, peer_note = 0, str(err)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            e.peer_close(dist.barrier)
            use_peer = False
            args.exchange = "nccl"
            peer_note = "peer window unavailable (%s): NCCL all-gather + index slices instead" % (peer_note or "a peer failed")
    if use_peer:
        wt = torch.as_tensor(Window(wptr, 3 * 16 * N), device=dev)
        pm4_regions = [wt[:4 * N], wt[4 * N:8 * N]]
        pm4_full = pm4_regions[0]
        acc_w, pot_w = wt[8 * N:11 * N], wt[11 * N:12 * N]
    else:
        pm4_full = torch.empty(N * 4, dtype=torch.float32, device=dev)
    xstate = {"n": 0}
    use_peer_at_start = use_peer
    if use_peer:
        acc, pot = acc_w, pot_w
    else:
        acc = torch.zeros(3 * N, dtype=torch.float32, device=dev)
        pot = torch.zeros(N, dtype=torch.float32, device=dev)
    slices, boxes = [], []
    for sim, box in snaps:
        pm = np.empty((N, 4), np.float32)
        pm[:, :3] = sim.pos.reshape(-1, 3); pm[:, 3] = sim.mas
        slices.append(torch.from_numpy(pm[i0:i1].copy()).to(dev).reshape(-1))
        boxes.append(np.ascontiguousarray(box, np.float32))
    import ctypes as C

    def gather(k):
        """Returns the tensor that holds all N records of snapshot k on this rank."""
        if use_peer:
            reg = xstate["n"] & 1; xstate["n"] += 1
            pm4_regions[reg][4 * i0:4 * i1].copy_(slices[k])
            if L.sfe_dev_peer_put(16 * (N * reg + i0), 16 * ni, sptr) != 0:
                raise SystemExit("bench.py: peer exchange failed: " + e.last_error())
            return pm4_regions[reg]
        if world > 1:
            dist.all_gather_into_tensor(pm4_full, slices[k])
        else:
            pm4_full.copy_(slices[k])      # at N=1 the "gather" is a device copy of the resident snapshot
        return pm4_full

    box_host = np.zeros(6, np.float32)

    def step_dev(k, ev=None):
        pm4 = gather(k)
        # root box of sim_solv() on the device (bit-identical to simulation.c:245-301; boxes[k] is the host mirror's) + build
        rc = L.sfe_dev_tree_build_box(pm4.data_ptr(), N, k, e.TREE_DEPTH_MAX, nodes_max, sptr)
        if ev is not None:
            ev[0].record(stream)
        if use_peer:
            # every rank walks a Morton-contiguous slot range and delivers acc / pot to the owners of the particles
            rc |= L.sfe_dev_bh_peer(pm4.data_ptr(), N, 16 * 2 * N, 16 * 2 * N + 12 * N, e.G, e.EPS2, THETA, flavour, sptr)
        else:
            rc |= L.sfe_dev_bh(pm4.data_ptr(), N, i0, ni, acc.data_ptr(), pot.data_ptr(), e.G, e.EPS2, THETA, flavour, sptr)
        if ev is not None:
            ev[1].record(stream)
        if rc != 0:
            raise SystemExit("bench.py: engine call failed: " + e.last_error())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up
    for w in range(max(args.warmup, 3)):
        step_dev(w % ring)
        if L.sfe_dev_last_box(box_host.ctypes.data, sptr) != 0 or not np.array_equal(box_host, boxes[w % ring]):
            raise SystemExit("bench.py: the device root box differs from the reference's")
    barrier()
    st = (C.c_uint32(), C.c_uint32(), C.c_uint32())
    if L.sfe_dev_tree_status(sptr, C.byref(st[0]), C.byref(st[1]), C.byref(st[2])) != 0:
        raise SystemExit("bench.py: tree build failed: " + e.last_error())

    # ---- timed region: K steps, device resident
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.wait_first()
        time.sleep(0.1)
    walk_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = L.sfe_launch_count()
    t_wall0 = time.time()
    e0.record(stream)
    for s in range(args.steps):
        step_dev((args.warmup + s) % ring, walk_ev[s])
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    launches = int(L.sfe_launch_count() - launches0)
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = N / (ms_step * 1e-3)
    walk_ms = float(np.mean([a.elapsed_time(b) for a, b in walk_ev]))
    walk_ms = max_over_ranks(walk_ms)

    # ---- the other walk flavour on the same steps, for information (FAST: same interaction lists, rsqrt.approx, 1e-6
    # rel-RMS from EXACT - inside the 1e-4 parity bar for Barnes-Hut; EXACT is bit-identical to the reference)
    other = e.WALK_FAST if flavour == e.WALK_EXACT else e.WALK_EXACT
    main_flavour, flavour = flavour, other
    for w in range(3):
        step_dev(w % ring)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(stream)
    for s in range(args.steps):
        step_dev((args.warmup + s) % ring)
    a1.record(stream)
    barrier()
    alt_ms = max_over_ranks(a0.elapsed_time(a1)) / args.steps
    flavour = main_flavour

    # ---- parity of the multi-GPU path on real devices: every rank's slice of one step against a single-rank walk of all
    # particles on rank 0's own tree (the EXACT walk is bit-reproducible), plus a checksum that is the same at every GPU count
    step_dev(0)
    barrier()
    parity = None
    mine_acc = acc[3 * i0:3 * i1].clone(); mine_pot = pot[i0:i1].clone()
    if world > 1 and N % world == 0:
        ga = [torch.empty_like(mine_acc) for _ in range(world)]; gp = [torch.empty_like(mine_pot) for _ in range(world)]
        dist.all_gather(ga, mine_acc); dist.all_gather(gp, mine_pot)
        if rank == 0:
            full_acc = torch.zeros(3 * N, dtype=torch.float32, device=dev); full_pot = torch.zeros(N, dtype=torch.float32, device=dev)
            pm4_now = pm4_regions[(xstate["n"] - 1) & 1] if use_peer else pm4_full
            if L.sfe_dev_bh(pm4_now.data_ptr(), N, 0, N, full_acc.data_ptr(), full_pot.data_ptr(), e.G, e.EPS2, THETA, flavour, sptr) != 0:
                raise SystemExit("bench.py: parity walk failed: " + e.last_error())
            torch.cuda.synchronize()
            same = bool(torch.equal(torch.cat(ga).view(torch.int32), full_acc.view(torch.int32)) and
                        torch.equal(torch.cat(gp).view(torch.int32), full_pot.view(torch.int32)))
            parity = {"bit_identical_to_single_rank": same, "flavour": args.flavour,
                      "checked": "acc and pot of all %d particles of one step, gathered from the %d ranks, against a walk of all "
                                 "particles on rank 0's own tree" % (N, world),
                      "checksum": "%016x" % slot_hash(torch, full_acc, full_pot)}
            del full_acc, full_pot
        del ga, gp
    elif world == 1:
        parity = {"bit_identical_to_single_rank": True, "flavour": args.flavour,
                  "checked": "single rank (the checksum is the one the multi-GPU runs must reproduce)",
                  "checksum": "%016x" % slot_hash(torch, acc, pot)}
    barrier()

    # ---- e2e: host buffers through the host C ABI (N=1) / pinned slices + collective + device ABI (N>1)
    e2e_ms = None
    h2d_b = d2h_b = 0
    if world == 1:
        for sim, box in snaps:
            L.sfe_host_register(sim.mem.ctypes.data, sim.mem.nbytes)
        def step_host(k):
            sim, box = snaps[k]
            rc = L.sfe_solv_tree(sim.pos.ctypes.data, sim.mas.ctypes.data, sim.acc.ctypes.data, sim.pot.ctypes.data, N, k,
                                 e.TREE_DEPTH_MAX, nodes_max, e.G, e.EPS2, THETA, flavour, box_host.ctypes.data)
            if rc != 0:
                raise SystemExit("bench.py: engine call failed: " + e.last_error())
        for w in range(3):
            step_host(w % ring)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(args.steps):
            step_host((args.warmup + s) % ring)
        torch.cuda.synchronize()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
        h2d_b, d2h_b = 16 * N, 16 * N
    else:
        host_in = [torch.from_numpy(np.ascontiguousarray(
            np.concatenate([snaps[k][0].pos.reshape(-1, 3)[i0:i1], snaps[k][0].mas[i0:i1, None]], axis=1))).pin_memory()
            for k in range(ring)]
        host_acc = torch.empty(3 * ni, dtype=torch.float32).pin_memory()
        host_pot = torch.empty(ni, dtype=torch.float32).pin_memory()
        def step_host(k):
            slices[k].copy_(host_in[k].reshape(-1), non_blocking=True)
            step_dev(k)
            host_acc.copy_(acc[3 * i0:3 * i1], non_blocking=True)
            host_pot.copy_(pot[i0:i1], non_blocking=True)
            stream.synchronize()
        for w in range(3):
            step_host(w % ring)
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            step_host((args.warmup + s) % ring)
        barrier()
        e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps)
        h2d_b, d2h_b = 16 * ni, 16 * ni
    if sampler:
        sampler.window(t_wall0, t_wall1)
        clocks = sampler.stop()
    else:
        clocks = None

    # ---- PP leg: interactions/s at the same N, target-sliced the same way
    pp = None
    if not args.no_pp:
        pm4 = gather(0)
        reps = 2 if N >= (1 << 19) else 5
        L.sfe_dev_pp(pm4.data_ptr(), N, i0, ni, acc.data_ptr(), pot.data_ptr(), e.G, e.EPS2, sptr)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for r in range(reps):
            pm4 = gather(r % ring)
            L.sfe_dev_pp(pm4.data_ptr(), N, i0, ni, acc.data_ptr(), pot.data_ptr(), e.G, e.EPS2, sptr)
        p1.record(stream)
        barrier()
        pp_ms = max_over_ranks(p0.elapsed_time(p1)) / reps
        info = e.device_info()
        inter = float(N) * float(N) / (pp_ms * 1e-3)
        tf = 20.0 * inter / 1e12
        # the same solve with unequal masses (the general kernel: the m_j multiply stays in the loop, 13 FMA-pipe operations per
        # interaction instead of 12); the reference's own ICs have equal masses, which is what `value` / `frac` are measured on
        pm4m = gather(0).clone()
        pm4m.view(-1, 4)[:, 3] *= torch.linspace(0.5, 1.5, N, device=dev)
        L.sfe_dev_pp(pm4m.data_ptr(), N, i0, min(ni, 65536), acc.data_ptr(), pot.data_ptr(), e.G, e.EPS2, sptr)
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record(stream)
        L.sfe_dev_pp(pm4m.data_ptr(), N, i0, ni, acc.data_ptr(), pot.data_ptr(), e.G, e.EPS2, sptr)
        q1.record(stream)
        barrier()
        mix_ms = max_over_ranks(q0.elapsed_time(q1))
        del pm4m
        tf_mix = 20.0 * float(N) * float(N) / (mix_ms * 1e-3) / 1e12
        pp = {"metric": "PP interactions/s", "value": inter, "ms_per_solve": pp_ms, "N": N,
              "flop_per_interaction": 20, "achieved_tflops": tf,
              "fp32_peak_tflops": info["fp32_peak_tflops"] * world,
              "peak_note": "2*128 lanes*%d SMs*%.0f MHz per GPU (cudaDevAttrClockRate)" % (info["sm_count"], info["clock_khz"] / 1e3),
              "frac": tf / (info["fp32_peak_tflops"] * world),
              "frac_equal_mass": tf / (info["fp32_peak_tflops"] * world),
              "frac_mixed": tf_mix / (info["fp32_peak_tflops"] * world), "ms_per_solve_mixed": mix_ms, "achieved_tflops_mixed": tf_mix,
              "note": "frac counts 20 flop per interaction in both cases; the equal-mass kernel (every IC the reference generates) executes 19"}

    # ---- C3 SPH kernels alone (single rank): density + EOS + pressure force on a resident snapshot
    sph = None
    if not args.no_sph and world == 1:
        sph = sph_leg(e, L, torch, dev, stream)

    # ---- vis_draw() replacement on the same state (single rank)
    vis = None
    if not args.no_vis and world == 1:
        vis = vis_leg(e, L, torch, dev, stream, N, gather(0), pot, not args.no_cpu)

    # ---- BASELINE.json's larger configurations: C5 at every GPU count (it fits one GPU), C4 where it is defined (2/4/8)
    c5 = c4 = None
    if not args.no_c5:
        c5 = c5_leg(e, L, torch, dist, world, rank, dev, stream, flavour, barrier, max_over_ranks)
    if not args.no_c4 and world > 1:
        c4 = c4_leg(e, L, torch, dist, world, rank, dev, stream, barrier, max_over_ranks)

    # ---- device-resident leapfrog: whole sim_step()s with the state in HBM, through sfe_sim_run()
    resident = None
    if not args.no_resident:
        barrier()
        if use_peer:
            del wt, pm4_regions, pm4_full, acc_w, pot_w, acc, pot
            e.peer_close(dist.barrier)
        resident = {}
        legs = [("C2", N, True, False)]
        if not args.no_sph:
            legs.append(("C3", 262144, False, True))
        for name, n_leg, tree, with_sph in legs:
            resident[name] = resident_leg(e, L, torch, dist, world, rank, n_leg, tree, with_sph, flavour, allgather_bytes,
                                          barrier, max_over_ranks, use_peer_at_start)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the walk): algorithmic bytes per launch / measured launch time
    hbm, hbm_src = measured_peaks()
    walk_bytes = (32.0 * V_C2 + 32.0) * ni            # 32 B per node visit + 32 B particle I/O
    step_bytes = (32.0 * V_C2 + 32.0 + 16.0 + 24.0 * 8 + 32.0 * 0.382) * N
    achieved = walk_bytes / (walk_ms * 1e-3) / 1e9
    traffic, winst, lanes, wavef, rec = None, None, None, None, {}
    info = e.device_info()
    # the engine walks a launch that fits the GPU in one wave with its latency kernel (launch_walk() in csrc/sfe_walk.cu)
    sp = ni <= info["sm_count"] * 7 * 128
    # ... and a launch whose gain covers the depth-first records with walk_p_kernel (the rule of launch_walk())
    p_pays = ni * (2.2e-4 if args.flavour == "fast" else 1.04e-4) > 1.3 * (30.0 + 12.0e-6 * N)
    kernel = "walk_sp_kernel" if sp else ("walk_p_kernel" if p_pays else "walk_kernel")   # (launch_walk() in csrc/sfe_walk.cu)
    prof = os.path.join(ROOT, "profiles", "walk_traffic.json")
    if os.path.exists(prof):
        try:
            rec = json.load(open(prof)).get(args.flavour + ("_sp" if sp else ("" if p_pays else "_walk_kernel")), {})
            share = ni / float(rec.get("targets", N_DEFAULT))       # the capture's counts, scaled to this launch's targets
            traffic = rec.get("dram_bytes_per_launch")
            winst = rec.get("warp_instructions_per_launch") * share
            wavef = rec.get("l1_data_pipe_wavefronts_per_launch") * share
            lanes = rec.get("lanes_per_instruction")
        except Exception:
            traffic = winst = wavef = None
    clock_hz = info["clock_khz"] * 1e3
    issue_peak = 4.0 * info["sm_count"] * clock_hz          # 4 schedulers per SM, one warp instruction per clock each
    l1_peak = 1.0 * info["sm_count"] * clock_hz             # one L1 data-pipe wavefront (128 B to the registers) per SM per clock
    hbm_model = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "peak_source": hbm_src,
                 "algorithmic_bytes_per_launch": walk_bytes,
                 "model": "SURVEY 8(d): 32 B x V(=%.1f node visits/particle, oracle-counted) + 32 B particle I/O, x %d targets" % (V_C2, ni),
                 "note": "not the binding roof: the node table is L1/L2 resident (DRAM traffic is 0.3 % of these bytes), so this "
                         "fraction exceeds 1; kept because SURVEY 8(d) and the north star quote the walk against HBM"}
    if winst and N == N_DEFAULT:
        issue = {"achieved": winst / (walk_ms * 1e-3), "peak": issue_peak, "unit": "warp-instructions/s",
                 "frac": winst / (walk_ms * 1e-3) / issue_peak}
        l1 = {"achieved": wavef / (walk_ms * 1e-3), "peak": l1_peak, "unit": "wavefronts/s", "frac": wavef / (walk_ms * 1e-3) / l1_peak,
              "model": "every 256-bit record load of a warp moves 1 KB into registers: >= 8 wavefronts at 128 B per clock per SM"}
        bound, top = ("issue", issue) if issue["frac"] >= l1["frac"] else ("l1", l1)
        roofline = {"kernel": "%s (sfe_dev_bh%s)" % (kernel, "_peer" if world > 1 else ""), "bound": bound, "achieved": top["achieved"],
                    "peak": top["peak"], "unit": top["unit"], "frac": top["frac"], "issue": issue, "l1_data_pipe": l1,
                    "traffic": traffic, "traffic_source": "profiles/walk_traffic.json: dram__bytes_read+write of one launch of this kernel "
                                                          "under ncu --set full (not measured inside this run)",
                    "warp_instructions_per_launch": winst, "l1_wavefronts_per_launch": wavef,
                    "instruction_source": "profiles/walk_traffic.json (%s): smsp__inst_executed.sum and l1tex__data_pipe_lsu_wavefronts "
                                          "of the same capture (%d targets), scaled to this launch's %d targets; peaks at the device's "
                                          "maximum SM clock" % (rec.get("source"), rec.get("targets", N_DEFAULT), ni),
                    "launch_ms": walk_ms, "lanes_active_per_instruction": lanes,
                    "model": "the walk is bound on the SM - instruction issue and the L1 data pipe (node records into registers), both "
                             "~70-80 % busy under ncu - not by memory: " + ("(N > 1: launch_ms includes the flag exchange)" if world > 1 else "single launch"),
                    "hbm_model": hbm_model}
    else:
        roofline = dict(hbm_model, kernel=kernel, traffic=traffic, launch_ms=walk_ms)
    roofline["whole_step"] = {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9,
                              "frac": step_bytes / (ms_step * 1e-3) / 1e9 / hbm,
                              "model": "SURVEY 8(d) B = 32 V + 32 + B_build = 16710 B per particle-step (HBM model, see hbm_model.note)"}
    # the tree build against its own HBM model (B_build = 16 + 24 D + 32 nodes/N bytes per particle)
    build_ms = ms_step - walk_ms
    build_bytes = (16.0 + 24.0 * int(st[2].value)) * N + 32.0 * int(st[0].value)
    roofline["tree_build"] = {"bound": "hbm", "ms": build_ms, "algorithmic_bytes": build_bytes, "achieved": build_bytes / (build_ms * 1e-3) / 1e9,
                              "peak": hbm, "unit": "GB/s", "frac": build_bytes / (build_ms * 1e-3) / 1e9 / hbm,
                              "model": "SURVEY 8(d): 16 + 24 D + 32 nodes/N bytes per particle; ms = step - walk (root box, exchange and build)"}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only, bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import sfo
        kind, threads, cores, times = cpu_reference_steps(N, 4, 1)
        cms = 1e3 * float(np.median(times))
        cpu = {"value": N / (cms * 1e-3), "unit": "particle-steps/s", "cores": threads, "host_cores": cores,
               "cpu_model": host_cpu_model(), "kind": kind, "build_flags": sfo.REF_BUILD_FLAGS,
               "sample": "4 full sim_solv() steps at N=%d after 1 warm-up (%.2f s each, median), %d OpenMP threads "
                         "(omp_get_max_threads) on %d host cores" % (N, cms / 1e3, threads, cores)}

    line = {
        "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(N),
                   "arm": "the engine on z-rotated copies of the reference IC, root box of step_number = ring position (the "
                          "reference arm re-solves the un-rotated IC with the root box of step_number = 0, 1, 2, ...: "
                          "statistically the same work)",
                   "N": N, "theta": THETA, "depth_max": 16, "walk": args.flavour,
                   "partition": ("tree per GPU; " + ("all-gather of {x,y,z,m} by the engine's peer-memory put + flag kernels "
                                 "(NVLink, sfe_dev_peer_put), walk by Morton slot range with results stored into the owner's window "
                                 "(sfe_dev_bh_peer)" if use_peer else "target index slices, all-gather of {x,y,z,m} by NCCL"))
                                if world > 1 else "single GPU",
                   "l2": "ring of %d distinct snapshots (%d MiB of inputs > L2) rotated between steps" % (ring, ring * 16 * N >> 20),
                   "tree_nodes": int(st[0].value), "tree_depth": int(st[2].value),
                   **({"note": peer_note} if peer_note else {})},
        "e2e": {"value": N / (e2e_ms * 1e-3), "unit": "particle-steps/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
                "api": "sfe_solv_tree (root box + tree_build + walk in one call) on pinned host arrays" if world == 1 else
                       "pinned slice H2D + exchange + sfe_dev_tree_build + walk + slice D2H"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "alt_walk": {"walk": "fast" if args.flavour == "exact" else "exact", "ms_per_step": alt_ms, "value": N / (alt_ms * 1e-3),
                     "unit": "particle-steps/s"},
        "parity": parity,
        "pp": pp,
        "c4": c4,
        "c5": c5,
        "sph": sph,
        "vis": vis,
        "resident": resident,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)
