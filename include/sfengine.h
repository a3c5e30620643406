/*
 * sfengine.h - C ABI of the B200-native force-evaluation engine for Starflood.
 *
 * This is the drop-in boundary: plain pointers and sizes, no reference structs, no
 * torch types.  The reference-side shims (starflood_b200/shim/) unpack sim_t / tree_t
 * and call these; INTEGRATION.md shows the Makefile change.  All file:line citations are
 * into the reference tree (/root/reference/src).
 *
 * Conventions
 *   - return 0 (STARFLOOD_SUCCESS, common.h:31) or -1 (STARFLOOD_FAILURE, common.h:32);
 *     diagnostics go to stderr as "<fn>() error: ..." like the reference (tree.c:181),
 *     and the text of the last one is available from sfe_last_error();
 *   - fp32 only (STARFLOOD_DOUBLE_PRECISION, config.h:221, is out of scope);
 *   - positions are xyz-interleaved (pos[3*i+k]) exactly as in sim.mem;
 *   - the *host* entry points are synchronous: inputs are read from and outputs are
 *     completely written to HOST memory when the call returns (simulation.c:423-451
 *     consumes acc/pot on the host right after sim_solv());
 *   - the *device* entry points (sfe_dev_*) take device pointers and a cudaStream_t
 *     passed as void*, enqueue work and return without synchronising.  All of them share the
 *     context's scratch buffers (reduction partials, cell lists, scan temporaries, the tree):
 *     issue them on ONE stream per context (any stream, but the same one for a whole sequence),
 *     and never concurrently with a host entry point;
 *   - there is NO CPU fallback: every entry point fails with -1 when no sm_100 device
 *     is usable.
 */
#ifndef SFENGINE_H
#define SFENGINE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFE_SUCCESS 0
#define SFE_FAILURE (-1)

/* Same memory layout as the reference's node_t (tree.h:18-24): 80 bytes. */
typedef struct {
    uint32_t parent;
    uint32_t bodies;
    uint32_t branch[8];
    float    bounds[6];
    float    params[4];
} sfe_node_t;

/* Per-phase device timings (CUDA events on the engine's stream) of the last host entry point. */
typedef struct {
    float h2d_ms;      /* host -> device copies + packing */
    float build_ms;    /* tree build (0 for PP / SPH) */
    float solve_ms;    /* force kernels */
    float d2h_ms;      /* device -> host copies */
    float total_ms;    /* first event to last event */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t launches; /* kernels launched by the call */
} sfe_timings_t;

/* Walk flavours for sfe_bh / sfe_dev_bh. */
#define SFE_WALK_EXACT 0u /* IEEE div/sqrt + the reference's Kahan order: bit-identical to gravity.c:273-660 */
#define SFE_WALK_FAST  1u /* rsqrt.approx, plain fp32 sums in the same visit order (<=1e-6 rel-RMS from EXACT) */

/* ---------------------------------------------------------------- lifecycle */
/* Bind the calling process to one CUDA device and create the engine context.  device < 0 picks
 * $SFE_DEVICE, else $LOCAL_RANK, else 0.  Idempotent.  Replaces nothing in the reference (it has no
 * accelerator runtime); the shims call it lazily from the first solver call. */
int sfe_init(int device);
int sfe_shutdown(void);
const char* sfe_last_error(void);
/* Pin a host range (e.g. the whole sim.mem block, simulation.c:89) so that the copies of the host entry
 * points run at full PCIe rate.  Optional: unpinned ranges are staged through an internal pinned buffer.
 * A registered range must be unregistered BEFORE it is freed (a stale registration makes later allocations
 * at the same address look pinned while the DMA engine still targets the old physical pages). */
int sfe_host_register(void* ptr, uint64_t bytes);
int sfe_host_unregister(void* ptr);
int sfe_get_timings(sfe_timings_t* out);
/* SM count, SM clock (kHz), and the FP32 FMA peak 2*128*SMs*clock in TFLOP/s of the bound device. */
int sfe_device_info(int* sm_count, int* clock_khz, double* fp32_peak_tflops);
/* Cumulative number of engine kernels launched by this process (host and device entry points). */
uint64_t sfe_launch_count(void);

/* ---------------------------------------------------------------- host entry points */
/* Replaces solve_gravity_part_part() (gravity.h:22, gravity.c:667-819): all-pairs softened gravity
 * including j == i; acc[3N] = -grav_g m_i sum(...), pot[N] = -grav_g sum(m_j / r).
 * (Parameter names avoid the reference's config.h macros G / EPSILON, which are visible where the shims include this.) */
int sfe_pp(const float* pos, const float* mas, float* acc, float* pot, uint32_t N, float grav_g, float eps2);
/* Same, for the target slice [i0, i0+ni) only; acc/pot are still indexed by global i. */
int sfe_pp_slice(const float* pos, const float* mas, float* acc, float* pot, uint32_t N,
                 uint32_t i0, uint32_t ni, float grav_g, float eps2);

/* Replaces tree_build() (tree.h:44, tree.c:74-318).  bounds = tree_ptr->bounds as set by sim_solv()
 * (simulation.c:298-301).  The tree stays on the device; nothing is written to host memory.
 * Fails like the reference on node overflow (tree.c:180-183) or depth overflow (tree.c:186-189). */
int sfe_tree_build(const float bounds[6], const float* pos, const float* mas, uint32_t N,
                   int depth_max, uint32_t nodes_max);
/* Replaces solve_gravity_part_tree() (gravity.h:14-18, gravity.c:245-663) using the tree of the last
 * sfe_tree_build() on the same pos/mas.  theta is TREE_THETA (config.h:44), compared against
 * width^2/dist^2 exactly as gravity.c:377,448.  The particle records are the ones sfe_tree_build() uploaded (the
 * reference calls build and walk back to back on the same arrays, simulation.c:312-321): pos / mas are not copied
 * again, and a host call that uploads other particles in between (sfe_pp, sfe_sph) retires the tree - sfe_bh() then
 * fails with "no tree" instead of walking a stale one. */
int sfe_bh(const float* pos, const float* mas, float* acc, float* pot, uint32_t N,
           float grav_g, float eps2, float theta, uint32_t flavour);
int sfe_bh_slice(const float* pos, const float* mas, float* acc, float* pot, uint32_t N,
                 uint32_t i0, uint32_t ni, float grav_g, float eps2, float theta, uint32_t flavour);
/* Test / debug hooks: copy the device tree out in the reference's node_t layout with canonical
 * (breadth-first) numbering, and the 48-bit path keys + sorted permutation of the last build. */
int sfe_tree_download(sfe_node_t* nodes, uint32_t max_nodes, uint32_t* num_nodes);
int sfe_tree_keys_download(uint64_t* keys, uint32_t* perm, uint32_t N, uint32_t* num_inbox);
/* Test hook: the EXACT walk evaluates 1/q, sqrt(q) and 1/sqrt(q) (gravity.c:754-759) with short Newton sequences
 * instead of the IEEE intrinsics; this compares the two for EVERY float of the range in which the sequences are
 * used and returns the number of mismatching bit patterns per function (must be {0, 0, 0}). */
int sfe_selftest_exact_math(uint64_t mismatches[3]);
/* Test / measurement hook: the walk has three kernels that produce the same bits - walk_kernel (throughput: launches that
 * fill the GPU), walk_sp_kernel (latency: a rank's share of the targets in a multi-GPU run, small N) and walk_p_kernel
 * (throughput on depth-first node records, one load per visit: what the FAST flavour uses for launches that fill the
 * GPU) - chosen from the number of targets and the flavour.  mode 0 / 1 / 2 forces the first / second / third for all
 * later walks of this process, -1 restores the choice. */
int sfe_walk_select(int mode);

/* The SOLVER_VIS part of solve_gravity_part_tree() (gravity.h:14-18 with SOLVER_VIS defined, the reference's default;
 * gravity.c:16-111, :380-445, :550-574): the debug overlay of particle `idx` (SOLVER_VIS = 3) - the
 * outline of every internal cell its walk looks at, coloured by depth, and a segment to every monopole it evaluates -
 * added (`+=`) to the host RGBA float buffer render_buffer[4*sizex*sizey] through the row-major 4x4 matrix mvp
 * (visualization.c's matMVP).  Uses the tree and the particle records of the last sfe_tree_build(); bit-identical to the
 * reference's drawing.  idx >= N draws nothing, like the reference. */
int sfe_bh_overlay(uint32_t idx, float theta, const float mvp[16], float* render_buffer, uint32_t sizex, uint32_t sizey);

/* Replaces vis_draw() (visualization.h:40, visualization.c:253-739): mean radius (:274-307), model-
 * view-projection matrix (:418-489), clear (:334-342), the particle splat - SPATIAL_SAMPLES Gaussian-jittered samples per
 * particle added to the integer accumulation buffer (:503-683) - and the finishing pass into the float RGBA render buffer
 * (:691-735).  The caller passes what visualization.c derives from config.h: mat_vm = matV * matM (row-major 4x4, :344-446),
 * ortho_scale = (real)exp2(ORTHO_SCALE), autoscale = AUTOSCALING defined, samples = SPATIAL_SAMPLES, pixel_value_scale =
 * 0.01 / SPATIAL_SAMPLES * exp2(EXPOSURE), tonemap = (0 < VIS_FILE_FORMAT).  Out: render_buffer[4*sizex*sizey] (host),
 * atomic_buffer[4*sizex*sizey] (host, may be NULL: nothing but vis_draw() itself reads it), average_radius (the value the
 * reference prints, may be NULL) and mvp[16] = the matMVP the reference leaves behind for the SOLVER_VIS overlay (may be
 * NULL).  The accumulation buffer is bit-identical to the reference's (integer adds commute; logf / sinf / cosf follow the
 * host C library bit for bit, csrc/sfe_libm.cuh); SHUTTER_SPEED (config.h:215, off) is not supported. */
int sfe_vis_draw(const float* pos, const float* pot, uint32_t N, uint32_t step_number, const float mat_vm[16],
                 float ortho_scale, int autoscale, uint32_t sizex, uint32_t sizey, uint32_t samples,
                 double pixel_value_scale, int tonemap, int32_t* atomic_buffer, float* render_buffer,
                 double* average_radius, float mvp[16]);

/* sim_solv()'s SOLVER_TREE branch in one call (simulation.c:245-325): root box of `step_number` (simulation.c:245-301,
 * returned in bounds_out when not NULL), tree_build (tree.c:74-318), solve_gravity_part_tree (gravity.c:245-663).
 * One upload of pos / mas, one download of acc / pot, no host synchronisation in between; a failed build is reported with
 * the reference's messages (tree.c:181,187); acc / pot are undefined then. */
int sfe_solv_tree(const float* pos, const float* mas, float* acc, float* pot, uint32_t N, uint32_t step_number,
                  int depth_max, uint32_t nodes_max, float grav_g, float eps2, float theta, uint32_t flavour,
                  float bounds_out[6]);

/* Replaces solve_sph() (sph.h:6, sph.c:14-283): density (all pairs, no cutoff), EOS
 * P = (eos_k rho)^(1/eos_n + 1), pressure force added to acc (read-modify-write) for r^2 <= cutoff2. */
int sfe_sph(const float* pos, const float* mas, const float* rad, float* rho, float* prs, float* acc,
            uint32_t N, float eos_k, float eos_n, float cutoff2);

/* ---------------------------------------------------------------- device entry points */
/* Particle records on the device are float4 {x, y, z, m} ("pm4"): the 16-byte unit that the
 * ncclAllGather of SURVEY 8(e) moves.  All pointers below are device pointers. */
/* sfe_vis_draw() on device buffers: pm4 {x,y,z,*}[N], pot[N]; atomic_buffer / render_buffer 16-byte aligned device arrays of
 * 4*sizex*sizey elements; info = device block of 16 floats (matMVP) followed by one double (average radius), 8-byte aligned. */
int sfe_dev_vis_draw(const void* pm4, const float* pot, uint32_t N, uint32_t step_number, const float mat_vm[16],
                     float ortho_scale, int autoscale, uint32_t sizex, uint32_t sizey, uint32_t samples,
                     double pixel_value_scale, int tonemap, int32_t* atomic_buffer, float* render_buffer, void* info,
                     void* stream);
int sfe_dev_pack_pm4(const float* pos, const float* mas, void* pm4, uint32_t n, void* stream);
int sfe_dev_pp(const void* pm4, uint32_t N, uint32_t i0, uint32_t ni, float* acc, float* pot,
               float grav_g, float eps2, void* stream);
/* Root box of sim_solv() for `step_number` (simulation.c:245-301) + tree_build, the box staying in device memory (no host
 * round trip between the two); sfe_dev_last_box() returns it once the stream has got there (it waits for the box only). */
int sfe_dev_tree_build_box(const void* pm4, uint32_t N, uint32_t step_number, int depth_max, uint32_t nodes_max, void* stream);
int sfe_dev_last_box(float bounds[6], void* stream);
int sfe_dev_tree_build(const float bounds[6], const void* pm4, uint32_t N, int depth_max,
                       uint32_t nodes_max, void* stream);
/* Raises the deferred build status (node / depth overflow) after synchronising the stream. */
int sfe_dev_tree_status(void* stream, uint32_t* num_nodes, uint32_t* num_inbox, uint32_t* depth);
int sfe_dev_bh(const void* pm4, uint32_t N, uint32_t i0, uint32_t ni, float* acc, float* pot,
               float grav_g, float eps2, float theta, uint32_t flavour, void* stream);
int sfe_dev_sph_density(const void* pm4, const float* rad, uint32_t N, uint32_t i0, uint32_t ni,
                        float* rho, void* stream);
/* Statistics hook of the SPH cell-list passes: (target, source) pairs evaluated since the last call (read and reset;
 * 0 when the all-pairs kernels did the work).  `stream` is the stream the passes were enqueued on (the call synchronises
 * it).  bench.py reports the count against the pairs inside the cutoff radius. */
int sfe_sph_pairs_evaluated(uint64_t* pairs, void* stream);
int sfe_dev_sph_eos(const float* rho, float* prs, uint32_t n, float eos_k, float eos_n, void* stream);
int sfe_dev_sph_force(const void* pm4, const float* rad, const float* rho, const float* prs, uint32_t N,
                      uint32_t i0, uint32_t ni, float cutoff2, float* acc, void* stream);


/* ---------------------------------------------------------------- resident leapfrog
 * The O(N) parts of sim_step() (simulation.c:348-519) and the root box of sim_solv() (simulation.c:245-301) as
 * device kernels, so that the state stays in HBM between steps.  vel / pos / ken / pen are bit-identical to the
 * reference's loops (no FMA contraction); the three double sums (mean radius, T, V) are exactly rounded sums where the
 * reference runs a serial Kahan fold (<= 2 ulp of double apart). */
/* kick_0 + drift (simulation.c:381-395) for the particles [i0, i0+ni): vel += 0.5 dt acc; pos += dt vel.
 * pm4 / vel / acc are indexed by global particle number. */
int sfe_dev_kick_drift(void* pm4, float* vel, const float* acc, uint32_t i0, uint32_t ni, float dt, void* stream);
/* Root box of sim_solv() (simulation.c:245-301): mean |pos| over all N particles, pcg4d(step_number) jitter.
 * `bounds` is HOST memory; the call returns when the device has published the box (it does not drain the stream). */
int sfe_dev_root_box(const void* pm4, uint32_t N, uint32_t step_number, float bounds[6], void* stream);
/* kick_1 (when do_kick != 0) + ken / pen + their sums (simulation.c:423-504) for [i0, i0+ni);
 * tv (device memory) receives {T, V} = {sum ken, sum pen} of the slice. */
int sfe_dev_kick_energy(const void* pm4, float* vel, const float* acc, const float* pot, float* ken, float* pen,
                        uint32_t i0, uint32_t ni, float dt, int do_kick, double* tv, void* stream);

/* Blocks of sim.mem (enum sim_conf, simulation.h:17-51) for sfe_sim_download(). */
#define SFE_SIM_POS (1u << 0)
#define SFE_SIM_VEL (1u << 1)
#define SFE_SIM_ACC (1u << 2)
#define SFE_SIM_MAS (1u << 3)
#define SFE_SIM_POT (1u << 4)
#define SFE_SIM_KEN (1u << 5)
#define SFE_SIM_PEN (1u << 6)
#define SFE_SIM_RAD (1u << 7)
#define SFE_SIM_RHO (1u << 8)
#define SFE_SIM_PRS (1u << 9)
#define SFE_SIM_ALL 0x3FFu
/* Solver selection = the reference's compile-time switches ENABLE_TREE / ENABLE_SPH (config.h:35,190). */
#define SFE_SOLVER_TREE 1u
#define SFE_SOLVER_SPH  2u
typedef struct {
    float    dt;          /* TIMESTEP_SIZE (config.h:170) */
    float    grav_g;      /* G (config.h:178) */
    float    eps2;        /* (real)(EPSILON*EPSILON) (gravity.c:754) */
    float    theta;       /* TREE_THETA (config.h:44) */
    int32_t  depth_max;   /* TREE_DEPTH_MAX (config.h:38) */
    uint32_t nodes_max;   /* TREE_NODES_MAX (config.h:41) */
    uint32_t flavour;     /* SFE_WALK_EXACT / SFE_WALK_FAST */
    uint32_t solver;      /* SFE_SOLVER_TREE | SFE_SOLVER_SPH; 0 = PP gravity only */
    float    eos_k, eos_n, cutoff2;   /* sph.c:148-149,231 */
} sfe_sim_params_t;
/* Copy a whole RAW-layout state block real[N*dof] (sim.mem, simulation.c:33-53; dof = SIM_DOF = 13 or 16) into HBM. */
int sfe_sim_upload(const float* mem, uint32_t N, uint32_t dof);
/* Copy the selected blocks back into a RAW-layout block (e.g. before sim_save_raw(), simulation_io.c:81-149). */
int sfe_sim_download(float* mem, uint32_t N, uint32_t dof, uint32_t blocks);
/* sim_solv() (simulation.c:231-345) on the resident state: root box + tree build + walk, or PP; then SPH. */
int sfe_sim_solv(const sfe_sim_params_t* p, uint32_t step_number);
/* sim_step() (simulation.c:348-519) on the resident state; energy = {E_tot, V, T} as written to statistics.csv
 * (simulation.c:499-503).  sfe_sim_run() does nsteps of them back to back (energy[3*nsteps]) with step numbers
 * first_step, first_step+1, ... (the reference's main loop sets sim.step_number before each step, main.c:451). */
int sfe_sim_step(const sfe_sim_params_t* p, uint32_t step_number, double energy[3]);
int sfe_sim_run(const sfe_sim_params_t* p, uint32_t first_step, uint32_t nsteps, double* energy);
/* sim_save_raw() (simulation_io.c:81-149) fed from the resident state: the same file - the whole
 * block real[N*SIM_DOF] - written WITHOUT stopping the leapfrog.  The call takes a device-side copy of the state on
 * the engine's stream (so the next step may start at once), a copy stream brings it to pinned host memory and a
 * writer thread puts it on disk.  Two snapshots may be in flight; a third call waits for the oldest.  In multi-rank
 * mode every rank writes the rows of its own target slice into the same file (all ranks must name the same path on a
 * file system they share, and call this collectively).  Errors of the writer (open / write / close) surface in
 * sfe_sim_save_wait(), which returns when every requested snapshot is on disk. */
int sfe_sim_save_raw_async(const char* filename);
int sfe_sim_save_wait(void);
/* The root box of the last tree solve (what sim_solv() would have put into sim.tree.bounds). */
/* vis_draw() of the resident state (single-rank mode): same as sfe_vis_draw() without the upload of pos / pot. */
int sfe_sim_vis_draw(uint32_t step_number, const float mat_vm[16], float ortho_scale, int autoscale, uint32_t sizex,
                     uint32_t sizey, uint32_t samples, double pixel_value_scale, int tonemap, int32_t* atomic_buffer,
                     float* render_buffer, double* average_radius, float mvp[16]);
int sfe_sim_last_box(float bounds[6]);
int sfe_sim_release(void);

/* ---------------------------------------------------------------- peer exchange, one process per GPU
 * The exchange step of the multi-GPU partition - "all-gather of {x,y,z,m} after each drift", and of rho between the
 * two SPH passes - done by the engine's own kernels over NVLink peer memory instead of a library collective: every
 * rank owns a WINDOW of device memory that all other ranks map through CUDA IPC, a producer kernel stores its slice
 * straight into every rank's window, and a one-warp kernel exchanges arrival flags (release/acquire at system scope).
 * No counterpart in the reference (it is single-node OpenMP).  The host only has to carry the 64-byte handles between
 * the processes once (any transport: torch.distributed, MPI, a file).
 *
 *   rank r:  sfe_peer_create(r, world, bytes, handle_r)   ->  exchange handles  ->  sfe_peer_open(handles[world])
 *
 * Exchanges are collective: every rank must enqueue the same sequence of them.  A region written by exchange k may
 * be overwritten by exchange k+2 at the earliest (ping-pong the regions), because a rank can run at most one
 * exchange ahead of its slowest peer.  A peer that never arrives makes the wait kernel give up after
 * $SFE_PEER_TIMEOUT_MS (default 10000) and the next host-synchronous entry point fail. */
#define SFE_PEER_MAX 16
#define SFE_PEER_HANDLE_BYTES 64
int sfe_peer_create(uint32_t rank, uint32_t world, uint64_t bytes, void* handle /* out: SFE_PEER_HANDLE_BYTES */);
int sfe_peer_open(const void* handles /* world x SFE_PEER_HANDLE_BYTES, indexed by rank */);
void* sfe_peer_window(void);   /* device pointer of this rank's window (NULL before sfe_peer_create) */
/* Store [offset, offset+bytes) of this rank's window into the same range of every peer's window, then exchange
 * arrival flags; work enqueued on `stream` after this call sees every rank's range.  bytes may differ per rank. */
int sfe_dev_peer_put(uint64_t offset, uint64_t bytes, void* stream);
/* Barnes-Hut walk on `world` ranks, collective.  Every rank has built the same tree from the same {x,y,z,m} array
 * (sfe_dev_tree_build).  Rank r walks the SLOTS [r*N/world, (r+1)*N/world) of the Morton-sorted order - spatial
 * neighbours, so its warps are as coherent as on one GPU, which a slice of particle indices is not - and stores acc / pot
 * of every particle it evaluated straight into the window of the rank that owns that particle's INDEX slice
 * [q*N/world, (q+1)*N/world), at acc_off + 12*i and pot_off + 4*i (byte offsets into the window, the same on all
 * ranks); the flag exchange follows.  On return of the stream every rank finds acc / pot of its own index slice in
 * its window; rows of other ranks' particles are not written. */
int sfe_dev_bh_peer(const void* pm4, uint32_t N, uint64_t acc_off, uint64_t pot_off, float grav_g, float eps2,
                    float theta, uint32_t flavour, void* stream);
/* Flags only (a barrier on the device, no host involvement). */
int sfe_dev_peer_barrier(void* stream);
/* Tear-down, in two phases so that no window is freed while a peer still maps it: every rank unmaps its peers'
 * windows (sfe_peer_detach), the HOSTS synchronise by their own means, then every rank frees its window
 * (sfe_peer_close; it detaches first when that has not happened yet).  A resident state that lives in the window
 * is released with it. */
int sfe_peer_detach(void);
int sfe_peer_close(void);
/* Window size that sfe_sim_upload() needs when the resident leapfrog runs on `world` ranks: two {x,y,z,m} arrays
 * (ping-pong between steps), acc and pot (the routed walk delivers them to the owner) and, for SIM_DOF = 16, rho.  With an open window of at least this size the sfe_sim_*
 * entry points partition the particles by target slice [rank*N/world, (rank+1)*N/world): kick_0 + drift stores the
 * moved slice into every rank's next {x,y,z,m} array (the drift and its all-gather are ONE kernel), every rank builds
 * the full tree and evaluates its slice, {T, V} are combined in rank order.  sfe_sim_download() then returns pos /
 * mas / rho / prs in full and the other blocks for the rank's slice only.  In this mode sfe_sim_upload(), sfe_sim_solv()
 * and sfe_sim_run() are collective (every rank calls them with the same arguments; upload ends with an exchange of
 * flags so that no rank stores into a window its owner is still filling). */
uint64_t sfe_sim_peer_bytes(uint32_t N, uint32_t dof);
/* The same plus room for the integer accumulation buffer of one sizex x sizey frame: with a window of at least this size
 * sfe_sim_vis_draw() works in multi-rank mode too (collective: every rank splats the particles of its slice into its
 * window, the buffers are summed over NVLink - integer adds, so every rank ends up with the single-rank frame bit for bit). */
uint64_t sfe_sim_peer_bytes_vis(uint32_t N, uint32_t dof, uint32_t sizex, uint32_t sizey);

#ifdef __cplusplus
}
#endif
#endif
