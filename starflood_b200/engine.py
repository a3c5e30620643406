"""Python host binding over the C ABI (include/sfengine.h) of libsfengine.so.

The product is the C ABI + the reference-side C shims (starflood_b200/shim/); this module exists so
that tests and bench.py can drive the same entry points, and it mirrors the reference's solver
interface for this path (same names, argument meaning and error behaviour):

    solve_gravity_part_part(sim)   gravity.h:22     tree_init / tree_free / tree_build   tree.h:38-44
    solve_gravity_part_tree(sim)   gravity.h:14-18  solve_sph(sim)                        sph.h:6

`Sim` is the reference's sim_t over one RAW-layout state block (simulation.h:17-67).
There is no CPU fallback: importing works without a GPU (so the symbol table can be checked), but
every compute call fails with SFE_FAILURE / raises when the CUDA library or device is missing.
"""
import ctypes as C
import os
import weakref

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFE_ENGINE_LIB") or os.path.join(HERE, "libsfengine.so")   # (the override: A/B builds of the library)

STARFLOOD_SUCCESS = 0   # common.h:31
STARFLOOD_FAILURE = -1  # common.h:32

# macro values the reference passes on this path (config.h:38-44,170-178; gravity.c:754; sph.c:148-149,231)
G = 1.000
EPSILON = 1.000e-3
EPS2 = np.float32(EPSILON * EPSILON)
TREE_THETA = 0.5
TREE_DEPTH_MAX = 16
TREE_NODES_MAX = 262144
EOS_K = 0.100
EOS_N = 1.000
SPH_CUTOFF2 = 1.000 * 1.000

TIMESTEP_SIZE = 1.0     # config.h:170
# vis_draw() (config.h:108,203-212): tanh tonemap for VIS_FILE_FORMAT > 0, SHUTTER_SPEED off
SPATIAL_SAMPLES = 128
EXPOSURE = -1.000
ORTHO_SCALE = -1.000
AUTOSCALING = True
VIS_FILE_FORMAT = 2

WALK_EXACT = 0
WALK_FAST = 1
SOLVER_TREE = 1         # ENABLE_TREE (config.h:35)
SOLVER_SPH = 2          # ENABLE_SPH (config.h:190)
SIM_BLOCKS = {"pos": 1 << 0, "vel": 1 << 1, "acc": 1 << 2, "mas": 1 << 3, "pot": 1 << 4, "ken": 1 << 5, "pen": 1 << 6,
              "rad": 1 << 7, "rho": 1 << 8, "prs": 1 << 9}
SIM_ALL = 0x3FF

NODE_DTYPE = np.dtype([
    ("parent", "<u4"), ("bodies", "<u4"), ("branch", "<u4", (8,)),
    ("bounds", "<f4", (6,)), ("params", "<f4", (4,)),
])


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("build_ms", C.c_float), ("solve_ms", C.c_float), ("d2h_ms", C.c_float),
                ("total_ms", C.c_float), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("launches", C.c_uint32)]


class SimParams(C.Structure):
    """sfe_sim_params_t: the compile-time configuration of the reference (config.h) as run-time values."""
    _fields_ = [("dt", C.c_float), ("grav_g", C.c_float), ("eps2", C.c_float), ("theta", C.c_float),
                ("depth_max", C.c_int32), ("nodes_max", C.c_uint32), ("flavour", C.c_uint32), ("solver", C.c_uint32),
                ("eos_k", C.c_float), ("eos_n", C.c_float), ("cutoff2", C.c_float)]


class EngineError(RuntimeError):
    pass


_P = C.c_void_p
_u32 = C.c_uint32
_f = C.c_float
SIGNATURES = {
    "sfe_init": ([C.c_int], C.c_int),
    "sfe_shutdown": ([], C.c_int),
    "sfe_last_error": ([], C.c_char_p),
    "sfe_host_register": ([_P, C.c_uint64], C.c_int),
    "sfe_host_unregister": ([_P], C.c_int),
    "sfe_get_timings": ([C.POINTER(Timings)], C.c_int),
    "sfe_device_info": ([C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)], C.c_int),
    "sfe_launch_count": ([], C.c_uint64),
    "sfe_pp": ([_P, _P, _P, _P, _u32, _f, _f], C.c_int),
    "sfe_pp_slice": ([_P, _P, _P, _P, _u32, _u32, _u32, _f, _f], C.c_int),
    "sfe_tree_build": ([_P, _P, _P, _u32, C.c_int, _u32], C.c_int),
    "sfe_bh": ([_P, _P, _P, _P, _u32, _f, _f, _f, _u32], C.c_int),
    "sfe_bh_slice": ([_P, _P, _P, _P, _u32, _u32, _u32, _f, _f, _f, _u32], C.c_int),
    "sfe_solv_tree": ([_P, _P, _P, _P, _u32, _u32, C.c_int, _u32, _f, _f, _f, _u32, _P], C.c_int),
    "sfe_tree_download": ([_P, _u32, C.POINTER(_u32)], C.c_int),
    "sfe_tree_keys_download": ([_P, _P, _u32, C.POINTER(_u32)], C.c_int),
    "sfe_selftest_exact_math": ([C.POINTER(C.c_uint64)], C.c_int),
    "sfe_walk_select": ([C.c_int], C.c_int),
    "sfe_sph": ([_P, _P, _P, _P, _P, _P, _u32, _f, _f, _f], C.c_int),
    "sfe_bh_overlay": ([_u32, _f, _P, _P, _u32, _u32], C.c_int),
    "sfe_vis_draw": ([_P, _P, _u32, _u32, _P, _f, C.c_int, _u32, _u32, _u32, C.c_double, C.c_int, _P, _P, _P, _P], C.c_int),
    "sfe_sim_vis_draw": ([_u32, _P, _f, C.c_int, _u32, _u32, _u32, C.c_double, C.c_int, _P, _P, _P, _P], C.c_int),
    "sfe_dev_vis_draw": ([_P, _P, _u32, _u32, _P, _f, C.c_int, _u32, _u32, _u32, C.c_double, C.c_int, _P, _P, _P, _P], C.c_int),
    "sfe_dev_pack_pm4": ([_P, _P, _P, _u32, _P], C.c_int),
    "sfe_dev_pp": ([_P, _u32, _u32, _u32, _P, _P, _f, _f, _P], C.c_int),
    "sfe_sph_pairs_evaluated": ([C.POINTER(C.c_uint64), _P], C.c_int),
    "sfe_dev_tree_build": ([_P, _P, _u32, C.c_int, _u32, _P], C.c_int),
    "sfe_dev_tree_build_box": ([_P, _u32, _u32, C.c_int, _u32, _P], C.c_int),
    "sfe_dev_last_box": ([_P, _P], C.c_int),
    "sfe_dev_tree_status": ([_P, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)], C.c_int),
    "sfe_dev_bh": ([_P, _u32, _u32, _u32, _P, _P, _f, _f, _f, _u32, _P], C.c_int),
    "sfe_dev_sph_density": ([_P, _P, _u32, _u32, _u32, _P, _P], C.c_int),
    "sfe_dev_sph_eos": ([_P, _P, _u32, _f, _f, _P], C.c_int),
    "sfe_dev_sph_force": ([_P, _P, _P, _P, _u32, _u32, _u32, _f, _P, _P], C.c_int),
    "sfe_dev_kick_drift": ([_P, _P, _P, _u32, _u32, _f, _P], C.c_int),
    "sfe_dev_root_box": ([_P, _u32, _u32, _P, _P], C.c_int),
    "sfe_dev_kick_energy": ([_P, _P, _P, _P, _P, _P, _u32, _u32, _f, C.c_int, _P, _P], C.c_int),
    "sfe_sim_upload": ([_P, _u32, _u32], C.c_int),
    "sfe_sim_download": ([_P, _u32, _u32, _u32], C.c_int),
    "sfe_sim_solv": ([_P, _u32], C.c_int),
    "sfe_sim_step": ([_P, _u32, _P], C.c_int),
    "sfe_sim_run": ([_P, _u32, _u32, _P], C.c_int),
    "sfe_sim_last_box": ([_P], C.c_int),
    "sfe_sim_release": ([], C.c_int),
    "sfe_sim_save_raw_async": ([C.c_char_p], C.c_int),
    "sfe_sim_save_wait": ([], C.c_int),
    "sfe_peer_create": ([_u32, _u32, C.c_uint64, _P], C.c_int),
    "sfe_peer_open": ([_P], C.c_int),
    "sfe_peer_window": ([], _P),
    "sfe_dev_peer_put": ([C.c_uint64, C.c_uint64, _P], C.c_int),
    "sfe_dev_peer_barrier": ([_P], C.c_int),
    "sfe_dev_bh_peer": ([_P, _u32, C.c_uint64, C.c_uint64, _f, _f, _f, _u32, _P], C.c_int),
    "sfe_peer_detach": ([], C.c_int),
    "sfe_peer_close": ([], C.c_int),
    "sfe_sim_peer_bytes": ([_u32, _u32], C.c_uint64),
    "sfe_sim_peer_bytes_vis": ([_u32, _u32, _u32, _u32], C.c_uint64),
}

_lib = None


def lib():
    """Load libsfengine.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError("libsfengine.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(or make -C starflood_b200); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def last_error():
    return (lib().sfe_last_error() or b"").decode()


def _check(rc, what):
    if rc != STARFLOOD_SUCCESS:
        raise EngineError("%s failed: %s" % (what, last_error()))


def init(device=-1):
    _check(lib().sfe_init(device), "sfe_init")


def shutdown():
    lib().sfe_shutdown()


def timings():
    t = Timings()
    lib().sfe_get_timings(C.byref(t))
    return t


def device_info():
    sm = C.c_int(); khz = C.c_int(); peak = C.c_double()
    _check(lib().sfe_device_info(C.byref(sm), C.byref(khz), C.byref(peak)), "sfe_device_info")
    return {"sm_count": sm.value, "clock_khz": khz.value, "fp32_peak_tflops": peak.value}


def _unregister(addr):
    if _lib is not None:
        _lib.sfe_host_unregister(addr)


def _ptr(a):
    return a.ctypes.data if a is not None else None


class Tree:
    """tree_t (tree.h:27-35).  `node` is host memory in the reference; the engine keeps the real tree on
    the device and fills `node` only on download()."""
    def __init__(self):
        self.max_nodes = 0
        self.num_nodes = 0
        self.bounds = np.zeros(6, np.float32)
        self.node = None


class Sim:
    """sim_t (simulation.h:53-67) over one RAW-layout block: real[N * SIM_DOF], sim_find(p) = mem + N*p."""
    OFFSETS = {"pos": (0, 3), "vel": (3, 3), "acc": (6, 3), "mas": (9, 1), "pot": (10, 1), "ken": (11, 1),
               "pen": (12, 1), "rad": (13, 1), "rho": (14, 1), "prs": (15, 1)}

    def __init__(self, N, sph=False, mem=None, pinned=False):
        self.N = int(N)
        self.step_number = 0
        self.sph = sph
        self.dof = 16 if sph else 13
        if mem is None:
            mem = np.zeros(self.N * self.dof, np.float32)
        assert mem.dtype == np.float32 and mem.size == self.N * self.dof and mem.flags.c_contiguous
        self.mem = mem
        self.tree = Tree()
        if pinned:
            # pin the block for full-rate PCIe copies; MUST be undone before the memory is freed, otherwise the
            # driver keeps a stale mapping and a later allocation at the same address is copied from / to the
            # wrong physical pages
            addr = self.mem.ctypes.data
            _check(lib().sfe_host_register(addr, self.mem.nbytes), "sfe_host_register")
            weakref.finalize(self, _unregister, addr)

    def find(self, name):
        first, width = self.OFFSETS[name]
        return self.mem[first * self.N:(first + width) * self.N]

    pos = property(lambda s: s.find("pos")); vel = property(lambda s: s.find("vel"))
    acc = property(lambda s: s.find("acc")); mas = property(lambda s: s.find("mas"))
    pot = property(lambda s: s.find("pot")); rad = property(lambda s: s.find("rad"))
    rho = property(lambda s: s.find("rho")); prs = property(lambda s: s.find("prs"))
    ken = property(lambda s: s.find("ken")); pen = property(lambda s: s.find("pen"))


# ---------------------------------------------------------------------------------- reference interface
def solve_gravity_part_part(sim, i0=0, ni=None):
    """gravity.h:22 / gravity.c:667-819.  Returns STARFLOOD_SUCCESS / STARFLOOD_FAILURE like the reference."""
    ni = sim.N - i0 if ni is None else ni
    return lib().sfe_pp_slice(_ptr(sim.pos), _ptr(sim.mas), _ptr(sim.acc), _ptr(sim.pot), sim.N, i0, ni,
                              G, EPS2)


def tree_init(tree, max_nodes):
    """tree.h:38 / tree.c:14-55."""
    tree.max_nodes = int(max_nodes)
    tree.num_nodes = 0
    tree.bounds[:] = 0
    tree.node = np.zeros(tree.max_nodes, NODE_DTYPE)
    return STARFLOOD_SUCCESS


def tree_free(tree):
    """tree.h:41 / tree.c:57-72."""
    tree.max_nodes = 0
    tree.num_nodes = 0
    tree.bounds[:] = 0
    tree.node = None
    return STARFLOOD_SUCCESS


def tree_build(tree, num_parts, pos, mas, depth_max=TREE_DEPTH_MAX):
    """tree.h:44 / tree.c:74-318.  tree.bounds must have been set by the caller (simulation.c:298-301)."""
    return lib().sfe_tree_build(_ptr(tree.bounds), _ptr(pos), _ptr(mas), num_parts, depth_max, tree.max_nodes)


def tree_download(tree):
    """Test hook: copy the device tree into tree.node (reference node_t layout, canonical numbering)."""
    n = _u32()
    rc = lib().sfe_tree_download(_ptr(tree.node), tree.max_nodes, C.byref(n))
    tree.num_nodes = n.value
    return rc


def tree_keys(N):
    keys = np.zeros(N, np.uint64); perm = np.zeros(N, np.uint32); nin = _u32()
    _check(lib().sfe_tree_keys_download(_ptr(keys), _ptr(perm), N, C.byref(nin)), "sfe_tree_keys_download")
    return keys, perm, nin.value


def selftest_exact_math():
    """Mismatch counts {1/q, sqrt(q), 1/sqrt(q)} of the EXACT walk's Newton sequences vs the IEEE intrinsics."""
    bad = (C.c_uint64 * 3)()
    _check(lib().sfe_selftest_exact_math(bad), "sfe_selftest_exact_math")
    return [int(b) for b in bad]


def walk_select(mode):
    """Test / measurement hook: 0 forces walk_kernel, 1 walk_sp_kernel, 2 walk_p_kernel (same bits, see sfengine.h), -1
    chooses by size and flavour."""
    _check(lib().sfe_walk_select(int(mode)), "sfe_walk_select")


def solve_gravity_part_tree(sim, flavour=WALK_EXACT, i0=0, ni=None, theta=TREE_THETA):
    """gravity.h:14-18 / gravity.c:245-663 (the SOLVER_VIS render-buffer arguments are ignored)."""
    ni = sim.N - i0 if ni is None else ni
    return lib().sfe_bh_slice(_ptr(sim.pos), _ptr(sim.mas), _ptr(sim.acc), _ptr(sim.pot), sim.N, i0, ni,
                              G, EPS2, theta, flavour)


def solv_tree(sim, flavour=WALK_EXACT, theta=TREE_THETA, depth_max=TREE_DEPTH_MAX):
    """The SOLVER_TREE branch of sim_solv() (simulation.c:245-325) as ONE engine call: root box of sim.step_number
    (left in sim.tree.bounds), tree_build, solve_gravity_part_tree."""
    return lib().sfe_solv_tree(_ptr(sim.pos), _ptr(sim.mas), _ptr(sim.acc), _ptr(sim.pot), sim.N, sim.step_number,
                               depth_max, sim.tree.max_nodes, G, EPS2, theta, flavour, _ptr(sim.tree.bounds))


SOLVER_VIS = 3   # config.h:32


def solve_gravity_part_tree_vis(sim, mvp, render_buffer, sizex, sizey, flavour=WALK_EXACT, idx=SOLVER_VIS, theta=TREE_THETA):
    """solve_gravity_part_tree(sim, render_buffer, sizex, sizey) of a SOLVER_VIS build (gravity.h:14-18): the walk, then
    the debug overlay of particle idx through the matrix the reference keeps in matMVP (visualization.c:251)."""
    rc = solve_gravity_part_tree(sim, flavour, theta=theta)
    if rc != 0:
        return rc
    m = np.ascontiguousarray(mvp, np.float32)
    return lib().sfe_bh_overlay(idx, theta, _ptr(m), _ptr(render_buffer), sizex, sizey)


class Vis:
    """vis_t (visualization.h:19-28): the integer accumulation buffer and the float RGBA render buffer."""
    def __init__(self):
        self.sizex = self.sizey = 0
        self.atomic_buffer = self.render_buffer = None
        self.average_radius = float("nan")
        self.matMVP = np.zeros(16, np.float32)     # visualization.c:251


def vis_init(vis, sizex, sizey):
    """visualization.h:31 / visualization.c:37-192 (host allocation only)."""
    vis.sizex, vis.sizey = int(sizex), int(sizey)
    vis.atomic_buffer = np.zeros(4 * vis.sizex * vis.sizey, np.int32)
    vis.render_buffer = np.zeros(4 * vis.sizex * vis.sizey, np.float32)
    return STARFLOOD_SUCCESS


def vis_free(vis):
    """visualization.h:34"""
    vis.sizex = vis.sizey = 0
    vis.atomic_buffer = vis.render_buffer = None
    return STARFLOOD_SUCCESS


_libm = None


def _f32_libm(name, x):
    """cosf / sinf of the host C library (numpy's float32 ufuncs are a different implementation)."""
    global _libm
    if _libm is None:
        _libm = C.CDLL("libm.so.6")
        for f in ("cosf", "sinf"):
            getattr(_libm, f).argtypes = [C.c_float]; getattr(_libm, f).restype = C.c_float
    return np.float32(getattr(_libm, name)(float(x)))


def vis_model_view(step_number, dt=TIMESTEP_SIZE):
    """matV * matM of vis_draw() (visualization.c:327,344-446): the model matrix is a slowly turning rotation, the view
    matrix the identity.  Same fp32 operations in the same order as the reference's C."""
    f = np.float32
    TAU = f(6.2831853071795864769)
    time = f(float(dt) * float(step_number))
    alp = f(0.0); bet = TAU * f(0.1250); gam = TAU * f(0.0625)
    alp = alp + TAU * f(1.000e-4) * time; bet = bet + TAU * f(1.000e-4) * time; gam = gam + TAU * f(1.000e-4) * time
    ca, sa = _f32_libm("cosf", alp), _f32_libm("sinf", alp)
    cb, sb = _f32_libm("cosf", bet), _f32_libm("sinf", bet)
    cg, sg = _f32_libm("cosf", gam), _f32_libm("sinf", gam)
    M = np.array([ca * cb, ca * sb * sg - sa * cg, ca * sb * cg + sa * sg, 0,
                  sa * cb, sa * sb * sg + ca * cg, sa * sb * cg - ca * sg, 0,
                  -sb, cb * sg, cb * cg, 0,
                  0, 0, 0, 1], np.float32)
    V = np.eye(4, dtype=np.float32).ravel()
    VM = np.zeros(16, np.float32)
    for i in range(4):
        for j in range(4):
            VM[4 * i + j] = (V[4 * i + 3] * M[12 + j]) + (V[4 * i + 2] * M[8 + j]) + (V[4 * i + 1] * M[4 + j]) + (V[4 * i] * M[j])
    return VM


def _vis_args():
    scale = 0.01 / float(SPATIAL_SAMPLES)
    if EXPOSURE is not None:
        scale *= 2.0 ** EXPOSURE
    ortho = np.float32(2.0 ** ORTHO_SCALE) if ORTHO_SCALE is not None else np.float32(1.0)
    return ortho, 1 if AUTOSCALING else 0, SPATIAL_SAMPLES, scale, 1 if VIS_FILE_FORMAT > 0 else 0


def vis_draw(vis, sim):
    """visualization.h:40 / visualization.c:253-739: renders sim into vis.atomic_buffer / vis.render_buffer and leaves
    matMVP behind like the reference does."""
    vm = vis_model_view(sim.step_number)
    ortho, auto, samples, scale, tonemap = _vis_args()
    rbar = C.c_double()
    rc = lib().sfe_vis_draw(_ptr(sim.pos), _ptr(sim.pot), sim.N, sim.step_number, _ptr(vm), ortho, auto, vis.sizex, vis.sizey,
                            samples, scale, tonemap, _ptr(vis.atomic_buffer), _ptr(vis.render_buffer), C.addressof(rbar),
                            _ptr(vis.matMVP))
    vis.average_radius = rbar.value
    return rc


def sim_vis_draw(vis, step_number):
    """vis_draw() of the HBM-resident state (sfe_sim_upload / sfe_sim_run): no particle upload."""
    vm = vis_model_view(step_number)
    ortho, auto, samples, scale, tonemap = _vis_args()
    rbar = C.c_double()
    rc = lib().sfe_sim_vis_draw(step_number, _ptr(vm), ortho, auto, vis.sizex, vis.sizey, samples, scale, tonemap,
                                _ptr(vis.atomic_buffer), _ptr(vis.render_buffer), C.addressof(rbar), _ptr(vis.matMVP))
    vis.average_radius = rbar.value
    return rc


def solve_sph(sim):
    """sph.h:6 / sph.c:14-283."""
    return lib().sfe_sph(_ptr(sim.pos), _ptr(sim.mas), _ptr(sim.rad), _ptr(sim.rho), _ptr(sim.prs), _ptr(sim.acc),
                         sim.N, EOS_K, EOS_N, SPH_CUTOFF2)


def root_box(sim):
    """Caller-side root box of sim_solv() (simulation.c:245-301): stays host code in the reference; this is
    the same arithmetic in numpy for hosts that do not link the reference (bench.py, tests)."""
    p = sim.pos.reshape(-1, 3).astype(np.float64)
    r = np.sqrt((p[:, 0] * p[:, 0]) + (p[:, 1] * p[:, 1]) + (p[:, 2] * p[:, 2]))
    # Kahan-compensated left fold == exactly rounded sum to well below fp32 resolution of the box
    import math
    rbar = math.fsum(r.tolist()) * (1.0 / sim.N)
    s = _pcg4d(np.array([0x638FE17B + sim.step_number, 0xB10BDCE7, 0x7B8ADB9B, 0xD8C5850E], np.uint64))
    inv = 2.3283062977608182109179324470460414886474609375e-10
    b = np.zeros(6, np.float32)
    for k in range(3):
        jitter = 2.0 * (inv * float(s[k])) - 1.0
        b[k] = np.float32(-2.0 * rbar + 0.5 * jitter * rbar)
        b[3 + k] = np.float32(2.0 * rbar + 0.5 * jitter * rbar)
    return b


def _pcg4d(v):
    """rng.h:13-40 on uint32 lanes."""
    M = np.uint64(0xFFFFFFFF)
    v = [(int(x) * 1664525 + 1013904223) & 0xFFFFFFFF for x in v]
    for rnd in range(2):
        v[0] = (v[0] + v[1] * v[3]) & 0xFFFFFFFF
        v[1] = (v[1] + v[2] * v[0]) & 0xFFFFFFFF
        v[2] = (v[2] + v[0] * v[1]) & 0xFFFFFFFF
        v[3] = (v[3] + v[1] * v[2]) & 0xFFFFFFFF
        if rnd == 0:
            v = [x ^ (x >> 16) for x in v]
    return v


def sim_solv(sim, tree=True, sph=False, flavour=WALK_EXACT):
    """The dispatch of sim_solv() (simulation.c:231-345) over the engine: root box -> tree_build ->
    solve_gravity_part_tree, or solve_gravity_part_part; then solve_sph."""
    if tree:
        sim.tree.bounds[:] = root_box(sim)
        tree_build(sim.tree, sim.N, sim.pos, sim.mas)      # return value ignored like simulation.c:312
        rc = solve_gravity_part_tree(sim, flavour)
    else:
        rc = solve_gravity_part_part(sim)
    if sph:
        rc = solve_sph(sim)
    return rc


# ---------------------------------------------------------------------------------- resident leapfrog
def sim_params(tree=True, sph=False, flavour=WALK_EXACT, dt=TIMESTEP_SIZE, nodes_max=TREE_NODES_MAX, theta=TREE_THETA,
               depth_max=TREE_DEPTH_MAX):
    return SimParams(dt, G, EPS2, theta, depth_max, nodes_max, flavour,
                     (SOLVER_TREE if tree else 0) | (SOLVER_SPH if sph else 0), EOS_K, EOS_N, SPH_CUTOFF2)


def sim_upload(sim):
    """Make the whole state block resident in HBM (sim.mem, simulation.c:33-53)."""
    return lib().sfe_sim_upload(_ptr(sim.mem), sim.N, sim.dof)


def sim_download(sim, blocks=SIM_ALL):
    """Copy blocks of the resident state back into sim.mem (names from Sim.OFFSETS or a bit mask)."""
    if not isinstance(blocks, int):
        m = 0
        for b in blocks:
            m |= SIM_BLOCKS[b]
        blocks = m
    return lib().sfe_sim_download(_ptr(sim.mem), sim.N, sim.dof, blocks)


def sim_solv_resident(sim, params):
    """sim_solv() (simulation.c:231-345) on the resident state, root box included."""
    return lib().sfe_sim_solv(C.byref(params), sim.step_number)


def sim_step(sim, params, nsteps=1):
    """sim_step() (simulation.c:348-519) nsteps times on the resident state, numbered sim.step_number, +1, ...
    Returns (rc, energy[nsteps, 3]) with the columns of statistics.csv: E_tot, V, T (simulation.c:499-503)."""
    en = np.zeros((nsteps, 3), np.float64)
    rc = lib().sfe_sim_run(C.byref(params), sim.step_number, nsteps, _ptr(en))
    return rc, en


def sim_save_raw_async(filename):
    """sim_save_raw() (simulation_io.c:81-149) from the resident state, overlapped with the following steps."""
    return lib().sfe_sim_save_raw_async(os.fsencode(filename))


def sim_save_wait():
    return lib().sfe_sim_save_wait()


def sim_last_box():
    b = np.zeros(6, np.float32)
    _check(lib().sfe_sim_last_box(_ptr(b)), "sfe_sim_last_box")
    return b


# ---------------------------------------------------------------------------------- peer exchange
PEER_HANDLE_BYTES = 64


def peer_setup(rank, world, nbytes, allgather):
    """Create this rank's exchange window, swap the IPC handles and map every peer's window (sfe_peer_create /
    sfe_peer_open).  `allgather(bytes) -> list of bytes indexed by rank` is the host-side transport, e.g.
    torch.distributed.all_gather_object; it is only used here, once.  Returns the window's device pointer."""
    h = C.create_string_buffer(PEER_HANDLE_BYTES)
    rc = lib().sfe_peer_create(rank, world, nbytes, h)
    why = last_error() if rc != STARFLOOD_SUCCESS else ""
    handles = allgather(h.raw if rc == STARFLOOD_SUCCESS else b"")       # always take part: the transport is collective
    if len(handles) != world or any(len(x) != PEER_HANDLE_BYTES for x in handles):
        lib().sfe_peer_close()
        raise EngineError("peer_setup: %s" % (why or "a peer could not create its window"))
    _check(lib().sfe_peer_open(b"".join(handles)), "sfe_peer_open")
    return lib().sfe_peer_window()


def peer_close(barrier=None):
    """Unmap the peers' windows, let the hosts meet (`barrier()`, e.g. torch.distributed.barrier), free this rank's
    window: no window disappears while a peer still maps it."""
    rc = lib().sfe_peer_detach()
    if barrier is not None:
        barrier()
    return rc | lib().sfe_peer_close()


def sim_slice(N, rank, world):
    """Target slice of a rank: [rank*N/world, (rank+1)*N/world)."""
    return (N * rank) // world, (N * (rank + 1)) // world
