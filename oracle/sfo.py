"""ctypes bindings for the CPU oracle (oracle/libsf_oracle.so) and, when present, the
unmodified reference solver libraries (oracle/_ref/libsfref_*.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by starflood_b200.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsf_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

NODE_DTYPE = np.dtype([
    ("parent", "<u4"), ("bodies", "<u4"), ("branch", "<u4", (8,)),
    ("bounds", "<f4", (6,)), ("params", "<f4", (4,)),
])
assert NODE_DTYPE.itemsize == 80


class WalkStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("iterations", "opened", "evals", "leaf_evals", "empty", "self_sub")]


def build(force=False):
    """Compile the oracle (and the reference libraries when /root/reference exists)."""
    src = os.path.join(HERE, "sf_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libsf_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(os.path.join(REF_DIR, "libsfref_pp.so"))):
        subprocess.check_call([os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)


_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_nodep = np.ctypeslib.ndpointer(dtype=NODE_DTYPE, flags="C_CONTIGUOUS")

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.sfo_pcg4d.argtypes = [_u32p]
        L.sfo_initcond.argtypes = [C.c_uint32, _f32p, _f32p, _f32p, C.c_void_p, C.c_double]
        L.sfo_initcond_galaxy_pair.argtypes = [C.c_uint32, _f32p, _f32p, _f32p]
        L.sfo_root_box.argtypes = [C.c_uint32, _f32p, C.c_uint32, _f32p]
        L.sfo_pp.argtypes = [C.c_uint32, _f32p, _f32p, C.c_float, C.c_float, C.c_uint32, C.c_uint32, _f32p, _f32p]
        L.sfo_pp_f64.argtypes = [C.c_uint32, _f32p, _f32p, C.c_double, C.c_double, C.c_uint32, C.c_uint32, _f64p, _f64p]
        L.sfo_tree_build_insert.argtypes = [_f32p, C.c_uint32, _f32p, _f32p, C.c_int, C.c_uint32, _nodep, _u32p]
        L.sfo_tree_build_insert.restype = C.c_int
        L.sfo_tree_build_sorted.argtypes = [_f32p, C.c_uint32, _f32p, _f32p, C.c_int, C.c_uint32, _nodep, _u32p,
                                            _u64p, _u32p, _u32p]
        L.sfo_tree_build_sorted.restype = C.c_int
        L.sfo_tree_canonicalise.argtypes = [_nodep, _nodep, C.c_uint32]
        L.sfo_tree_canonicalise.restype = C.c_uint32
        L.sfo_bh_walk.argtypes = [_nodep, C.c_uint32, _f32p, _f32p, C.c_float, C.c_float, C.c_float, C.c_int,
                                  C.c_uint32, C.c_uint32, _f32p, _f32p, C.POINTER(WalkStats)]
        L.sfo_sph_density.argtypes = [C.c_uint32, _f32p, _f32p, _f32p, C.c_uint32, C.c_uint32, _f32p]
        L.sfo_sph_eos.argtypes = [C.c_uint32, _f32p, C.c_float, C.c_float, _f32p]
        L.sfo_sph_force.argtypes = [C.c_uint32, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float,
                                    C.c_uint32, C.c_uint32, _f32p]
        L.sfo_kick.argtypes = [C.c_uint32, _f32p, _f32p, C.c_float]
        L.sfo_drift.argtypes = [C.c_uint32, _f32p, _f32p, C.c_float]
        L.sfo_energies.argtypes = [C.c_uint32, _f32p, _f32p, _f32p, _f32p, _f32p,
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
        L.sfo_vis_mean_radius.argtypes = [C.c_uint32, _f32p]
        L.sfo_vis_mean_radius.restype = C.c_double
        L.sfo_vis_matrices.argtypes = [C.c_double, C.c_uint32, C.c_double, C.c_double, C.c_int, _f32p, _f32p]
        L.sfo_vis_splat.argtypes = [C.c_uint32, _f32p, _f32p, C.c_uint32, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _i32p]
        L.sfo_vis_finish.argtypes = [_i32p, C.c_uint32, C.c_uint32, C.c_double, C.c_int, _f32p]
        _lib = L
    return _lib


# Reference macro values the shim passes (config.h:170-178, :38-44; gravity.c:754; sph.c:148-149,231)
G = np.float32(1.0)
EPS2 = np.float32(1.000e-3 * 1.000e-3)
THETA = np.float32(0.5)
DEPTH_MAX = 16
EOS_K = np.float32(0.100)
EOS_N = np.float32(1.000)
SPH_CUTOFF2 = np.float32(1.000 * 1.000)


def pcg4d(s):
    s = np.ascontiguousarray(s, dtype=np.uint32).copy()
    lib().sfo_pcg4d(s)
    return s


def initcond(N, outflow=4.0e-3, sph=False):
    pos = np.zeros(3 * N, np.float32); vel = np.zeros(3 * N, np.float32); mas = np.zeros(N, np.float32)
    rad = np.zeros(N, np.float32) if sph else None
    lib().sfo_initcond(N, pos, vel, mas, rad.ctypes.data if sph else None, outflow)
    return (pos, vel, mas, rad) if sph else (pos, vel, mas)


def initcond_galaxy_pair(N):
    pos = np.zeros(3 * N, np.float32); vel = np.zeros(3 * N, np.float32); mas = np.zeros(N, np.float32)
    lib().sfo_initcond_galaxy_pair(N, pos, vel, mas)
    return pos, vel, mas


def root_box(pos, step_number=0):
    b = np.zeros(6, np.float32)
    lib().sfo_root_box(len(pos) // 3, pos, step_number, b)
    return b


def pp(pos, mas, i0=0, i1=None, G_=G, eps2=EPS2):
    N = len(mas); i1 = N if i1 is None else i1
    acc = np.zeros(3 * N, np.float32); pot = np.zeros(N, np.float32)
    lib().sfo_pp(N, pos, mas, G_, eps2, i0, i1, acc, pot)
    return acc, pot


def pp_f64(pos, mas, i0=0, i1=None, G_=1.0, eps2=float(EPS2)):
    N = len(mas); i1 = N if i1 is None else i1
    acc = np.zeros(3 * N, np.float64); pot = np.zeros(N, np.float64)
    lib().sfo_pp_f64(N, pos, mas, G_, eps2, i0, i1, acc, pot)
    return acc, pot


def tree_build_insert(bounds, pos, mas, max_nodes, depth_max=DEPTH_MAX):
    N = len(mas)
    nodes = np.zeros(max_nodes, NODE_DTYPE); n = np.zeros(1, np.uint32)
    rc = lib().sfo_tree_build_insert(np.ascontiguousarray(bounds, np.float32), N, pos, mas, depth_max, max_nodes, nodes, n)
    return rc, nodes, int(n[0])


def tree_build_sorted(bounds, pos, mas, max_nodes, depth_max=DEPTH_MAX):
    N = len(mas)
    nodes = np.zeros(max_nodes, NODE_DTYPE); n = np.zeros(1, np.uint32)
    keys = np.zeros(N, np.uint64); perm = np.zeros(N, np.uint32); nin = np.zeros(1, np.uint32)
    rc = lib().sfo_tree_build_sorted(np.ascontiguousarray(bounds, np.float32), N, pos, mas, depth_max, max_nodes,
                                     nodes, n, keys, perm, nin)
    return rc, nodes, int(n[0]), keys, perm, int(nin[0])


def tree_canonicalise(nodes, max_out=None):
    max_out = len(nodes) if max_out is None else max_out
    out = np.zeros(max_out, NODE_DTYPE)
    n = lib().sfo_tree_canonicalise(np.ascontiguousarray(nodes), out, max_out)
    return out, int(n)


def bh_walk(nodes, pos, mas, i0=0, i1=None, G_=G, eps2=EPS2, theta=THETA, depth_max=DEPTH_MAX):
    N = len(mas); i1 = N if i1 is None else i1
    acc = np.zeros(3 * N, np.float32); pot = np.zeros(N, np.float32); st = WalkStats()
    lib().sfo_bh_walk(np.ascontiguousarray(nodes), N, pos, mas, G_, eps2, theta, depth_max, i0, i1, acc, pot, C.byref(st))
    return acc, pot, st


def bh_overlay(nodes, pos, mas, idx, mvp, rb, sizex, sizey, theta=THETA, depth_max=DEPTH_MAX):
    """SOLVER_VIS overlay of particle idx added into rb (float32[sizey*sizex*4]); gravity.c:16-111,380-445,550-574."""
    L = lib()
    L.sfo_bh_overlay.argtypes = [C.c_void_p, C.c_uint32, _f32p, _f32p, C.c_float, C.c_int, C.c_uint32, _f32p, _f32p,
                                 C.c_uint32, C.c_uint32]
    L.sfo_bh_overlay.restype = None
    nodes = np.ascontiguousarray(nodes)
    L.sfo_bh_overlay(nodes.ctypes.data, len(mas), pos, mas, theta, depth_max, idx, np.ascontiguousarray(mvp, np.float32), rb,
                     sizex, sizey)
    return rb


# vis_draw() parameters of the shipped configuration (config.h:170,203-212; VIS_FILE_FORMAT 2 -> tanh tonemap)
VIS_SAMPLES = 128
VIS_DT = 1.0
VIS_ORTHO_EXP2 = -1.0
VIS_AUTOSCALE = 1
VIS_EXPOSURE = -1.0
VIS_TONEMAP = 1


def vis_pixel_scale(samples=VIS_SAMPLES, exposure=VIS_EXPOSURE):
    """visualization.c:691-695"""
    scale = 0.01 / float(samples)
    if exposure is not None:
        scale *= 2.0 ** exposure
    return scale


def vis_mean_radius(pos):
    """visualization.c:274-307"""
    return float(lib().sfo_vis_mean_radius(len(pos) // 3, pos))


def vis_matrices(average_radius, step_number, dt=VIS_DT, ortho_exp2=VIS_ORTHO_EXP2, autoscale=VIS_AUTOSCALE):
    """visualization.c:327,344-489 -> (matV * matM, matMVP), row-major 4x4"""
    vm = np.zeros(16, np.float32); mvp = np.zeros(16, np.float32)
    lib().sfo_vis_matrices(average_radius, step_number, dt, ortho_exp2, autoscale, vm, mvp)
    return vm, mvp


def vis_draw(pos, pot, step_number, sizex, sizey, samples=VIS_SAMPLES, exposure=VIS_EXPOSURE, tonemap=VIS_TONEMAP,
             dt=VIS_DT, ortho_exp2=VIS_ORTHO_EXP2, autoscale=VIS_AUTOSCALE):
    """The whole vis_draw() (visualization.c:253-739) -> (average_radius, matMVP, atomic_buffer, render_buffer)."""
    N = len(pot)
    rbar = vis_mean_radius(pos)
    _, mvp = vis_matrices(rbar, step_number, dt, ortho_exp2, autoscale)
    ab = np.zeros(4 * sizex * sizey, np.int32)
    rb = np.zeros(4 * sizex * sizey, np.float32)
    lib().sfo_vis_splat(N, pos, pot, step_number, mvp, sizex, sizey, samples, ab)
    lib().sfo_vis_finish(ab, sizex, sizey, vis_pixel_scale(samples, exposure), tonemap, rb)
    return rbar, mvp, ab, rb


def sph_density(pos, mas, rad, i0=0, i1=None):
    N = len(mas); i1 = N if i1 is None else i1
    rho = np.zeros(N, np.float32)
    lib().sfo_sph_density(N, pos, mas, rad, i0, i1, rho)
    return rho


def sph_eos(rho, eos_k=EOS_K, eos_n=EOS_N):
    prs = np.zeros_like(rho)
    lib().sfo_sph_eos(len(rho), rho, eos_k, eos_n, prs)
    return prs


def sph_force(pos, mas, rad, rho, prs, acc, i0=0, i1=None, cutoff2=SPH_CUTOFF2):
    N = len(mas); i1 = N if i1 is None else i1
    lib().sfo_sph_force(N, pos, mas, rad, rho, prs, cutoff2, i0, i1, acc)
    return acc


def energies(vel, mas, pot):
    N = len(mas)
    ken = np.zeros(N, np.float32); pen = np.zeros(N, np.float32)
    T = C.c_double(); V = C.c_double()
    lib().sfo_energies(N, vel, mas, pot, ken, pen, C.byref(T), C.byref(V))
    return T.value, V.value


# ----------------------------------------------------------------------------------------------
# The unmodified reference, compiled by oracle/build_ref.sh
# ----------------------------------------------------------------------------------------------
class RefTree(C.Structure):          # tree.h:27-35
    _fields_ = [("max_nodes", C.c_uint32), ("num_nodes", C.c_uint32),
                ("bounds", C.c_float * 6), ("node", C.c_void_p)]


class RefSimTree(C.Structure):       # simulation.h:53-67 with ENABLE_TREE
    _fields_ = [("step_number", C.c_uint), ("N", C.c_uint), ("mem", C.c_void_p), ("tree", RefTree)]


class RefSimFlat(C.Structure):       # simulation.h:53-67 without ENABLE_TREE
    _fields_ = [("step_number", C.c_uint), ("N", C.c_uint), ("mem", C.c_void_p)]


REF_NODES_MAX = {"tree": 1048576, "tree_big": 8388608, "tree_vis": 1048576, "vis": 1048576}


class RefVis(C.Structure):           # visualization.h:19-28
    _fields_ = [("sizex", C.c_uint), ("sizey", C.c_uint), ("mem", C.c_void_p), ("atomic_buffer", C.c_void_p),
                ("render_buffer", C.c_void_p), ("binary_buffer", C.c_void_p)]


def ref_available(variant="pp"):
    return os.path.exists(os.path.join(REF_DIR, "libsfref_%s.so" % variant))


# Flags oracle/build_ref.sh compiles the reference with (reported next to every CPU number, bench.py cpu_baseline)
REF_BUILD_FLAGS = "-fopenmp -O3 -march=x86-64-v3 -flto=auto -std=c99"


def omp_threads(n=None):
    """Thread count of the OpenMP runtime the reference libraries run on (libgomp.so.1, the one libsfref_*.so link).
    With n: omp_set_num_threads(n) first.  An environment variable is not enough: a launcher may have exported
    OMP_NUM_THREADS=1 (torch.distributed.run does) and libgomp reads it only once, when it is loaded - so the count is
    set and read back through the runtime's own API (README.md:242-248 of the reference: OMP_NUM_THREADS=$(nproc))."""
    g = C.CDLL("libgomp.so.1")
    if n is not None:
        g.omp_set_num_threads(int(n))
    g.omp_get_max_threads.restype = C.c_int
    return int(g.omp_get_max_threads())


def host_cores():
    """Host cores this process may run on (what $(nproc) prints)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class quiet_stdout:
    """The reference printf()s from inside the solvers; silence fd 1 around calls."""
    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)
    def __exit__(self, *a):
        C.CDLL(None).fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._null); os.close(self._saved)


class RefSim:
    """A reference sim_t over a numpy-owned state block (RAW snapshot layout)."""
    def __init__(self, variant, N):
        self.variant = variant
        self.lib = C.CDLL(os.path.join(REF_DIR, "libsfref_%s.so" % variant))
        self.has_tree = variant.startswith("tree") or variant == "vis"
        self.dof = 16 if variant.endswith("sph") else 13
        self.N = N
        self.mem = np.zeros(N * self.dof, np.float32)
        self.sim = (RefSimTree if self.has_tree else RefSimFlat)()
        self.sim.step_number = 0; self.sim.N = N; self.sim.mem = self.mem.ctypes.data
        if self.has_tree:
            self.nodes_max = REF_NODES_MAX[variant]
            self.lib.tree_init.argtypes = [C.POINTER(RefTree), C.c_uint]
            assert self.lib.tree_init(C.byref(self.sim.tree), self.nodes_max) == 0
        for f in ("sim_solv", "sim_step", "initcond_generate", "solve_gravity_part_part"):
            getattr(self.lib, f).argtypes = [C.c_void_p]
            getattr(self.lib, f).restype = C.c_int

    def block(self, first, width=1):
        return self.mem[first * self.N:(first + width) * self.N]
    pos = property(lambda s: s.block(0, 3)); vel = property(lambda s: s.block(3, 3))
    acc = property(lambda s: s.block(6, 3)); mas = property(lambda s: s.block(9))
    pot = property(lambda s: s.block(10)); ken = property(lambda s: s.block(11)); pen = property(lambda s: s.block(12))
    rad = property(lambda s: s.block(13)); rho = property(lambda s: s.block(14)); prs = property(lambda s: s.block(15))

    def initcond(self):
        with quiet_stdout():
            self.lib.initcond_generate(C.byref(self.sim))

    def solv(self, step_number=0):
        self.sim.step_number = step_number
        with quiet_stdout():
            self.lib.sim_solv(C.byref(self.sim))

    def step(self, step_number):
        self.sim.step_number = step_number
        with quiet_stdout():
            self.lib.sim_step(C.byref(self.sim))

    def solv_vis(self, step_number, mvp, rb, sizex, sizey):
        """sim_solv() of the SOLVER_VIS build (variant tree_vis): simulation.c:231-232 takes the render buffer."""
        assert self.variant == "tree_vis"
        m = (C.c_float * 16).in_dll(self.lib, "matMVP")
        m[:] = [float(x) for x in np.asarray(mvp, np.float32).ravel()]
        self.sim.step_number = step_number
        self.lib.sim_solv.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint]
        with quiet_stdout():
            self.lib.sim_solv(C.byref(self.sim), rb.ctypes.data, sizex, sizey)

    def vis_draw(self, step_number, sizex, sizey):
        """The reference's own vis_init() + vis_draw() + vis_free() (variant vis, visualization.c:37-739) on the
        current state -> (matMVP, atomic_buffer, render_buffer) copies."""
        assert self.variant == "vis"
        vis = RefVis()
        L = self.lib
        L.vis_init.argtypes = [C.POINTER(RefVis), C.c_uint, C.c_uint]
        L.vis_draw.argtypes = [C.POINTER(RefVis), C.c_void_p]
        L.vis_free.argtypes = [C.POINTER(RefVis)]
        self.sim.step_number = step_number
        with quiet_stdout():
            assert L.vis_init(C.byref(vis), sizex, sizey) == 0
            assert L.vis_draw(C.byref(vis), C.byref(self.sim)) == 0
        n = 4 * sizex * sizey
        ab = np.frombuffer((C.c_int32 * n).from_address(vis.atomic_buffer), np.int32).copy()
        rb = np.frombuffer((C.c_float * n).from_address(vis.render_buffer), np.float32).copy()
        mvp = np.array((C.c_float * 16).in_dll(L, "matMVP")[:], np.float32)
        with quiet_stdout():
            L.vis_free(C.byref(vis))
        return mvp, ab, rb

    def nodes(self):
        """The reference's node table after the last tree_build (num_nodes is never stored, tree.c:77;
        it is 1 + 8 * #internal)."""
        buf = (C.c_char * (80 * self.nodes_max)).from_address(self.sim.tree.node)
        allnodes = np.frombuffer(buf, dtype=NODE_DTYPE)
        n = 1 + 8 * int((allnodes["bodies"] > 8).sum())
        return allnodes[:n]

    def close(self):
        if self.has_tree and self.sim.tree.node:
            self.lib.tree_free.argtypes = [C.POINTER(RefTree)]
            self.lib.tree_free(C.byref(self.sim.tree))
