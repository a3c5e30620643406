// sfe_internal.cuh - engine context, error plumbing and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include "../../include/sfengine.h"

namespace sfe {

// ---------------------------------------------------------------------------- errors
void set_error(const char* fn, const char* fmt, ...);
#define SFE_CUDA(fn, call)                                                              \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            sfe::set_error(fn, "%s failed: %s", #call, cudaGetErrorString(e__));        \
            return SFE_FAILURE;                                                         \
        }                                                                               \
    } while (0)

// ---------------------------------------------------------------------------- growable device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes, const char* fn) {
        if (bytes <= cap) return SFE_SUCCESS;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { set_error(fn, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); return SFE_FAILURE; }
        cap = want;
        return SFE_SUCCESS;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// ---------------------------------------------------------------------------- tree (host-side handle)
constexpr int SFE_SIDE_STREAMS = 4;  // the moment folds of consecutive levels overlap with each other and with the partitioning
constexpr int SFE_MAX_LEVELS = 21;   // deepest supported TREE_DEPTH_MAX (21 * 3 = 63 key bits; the reference uses 16)

struct Tree {
    // configuration of the last build
    uint32_t N = 0, nodes_max = 0;
    int depth_max = 0;
    DevBuf bounds_buf;            // float[6]  the root box of a build whose caller passed it in host memory
    uint32_t level0_range_len = 0;       // non-zero: the last build took the one-pass top levels and left srec[0] unwritten (top_level0_kernel)
    const float* bounds_dev = nullptr;   // the root box of the last build (bounds_buf or the root-box kernel's output)
    bool valid = false;
    // what the top-level pass of the last build told the host (mapped memory, read while the host waits for that pass anyway):
    // valid only for that build; lets tree_root_moments() skip the root gate's nine launches when no target can accept the root
    bool gate_hint = false; uint32_t hint_inbox = 0; float hint_kr = 0.f;
    bool recp_ready = false;   // the pre-order walk records (WalkRecP) of this build exist (made on demand: tree_ensure_recp)
    // results, fetched lazily by tree_status()
    uint32_t num_nodes = 0, num_inbox = 0, depth = 0;
    int status = 0;  // 0 ok, 1 node overflow, 2 depth overflow
    bool status_fetched = false;
    const void* pm4_of_build = nullptr;
    int final_perm = 0;
    // device storage
    DevBuf nodes;        // sfe_node_t[nodes_max]  reference layout, canonical (breadth-first) numbering
    DevBuf seg;          // uint2[nodes_max]       slot range of each node
    DevBuf next;         // uint32_t[nodes_max]    depth-first successor when a node is not opened
    DevBuf child_base;   // uint32_t[nodes_max]    first child of internal nodes
    DevBuf counts;       // uint32_t[8*nodes_max]  per-octant member counts of internal nodes
    DevBuf keys;         // uint64_t[N]            48-bit path keys by particle index (on demand)
    DevBuf srec[SFE_MAX_LEVELS + 3];   // SlotRec[N]  the particle order of every level: {x,y,z,m | key | index | owning node}
    DevBuf perm_out;     // uint32_t[N]            out-of-box particles in ascending index
    DevBuf oct;          // uint8_t[N]
    DevBuf rank;         // uint32_t[N]            stable rank inside (node, octant)
    DevBuf nmid;         // float4[nodes_max]      split planes of internal nodes
    DevBuf lookback;     // tickets + look-back descriptors + long-node lists of one build (zeroed per build)
    DevBuf xl_sum, xl_map, xl_flag, xl_list;   // scratch of the exact chunked fold
    DevBuf sub_list;         // uint2[N / 9 + 1]  subtree roots (sfe_subtree.cuh)
    DevBuf top_table, top_cols, top_rank;   // one-pass construction of the upper levels (sfe_top.cuh)
    uint32_t* summary_host = nullptr;  // mapped pinned: per-level summaries written by the device
    uint32_t* summary_dev = nullptr;
    uint32_t build_id = 0;
    DevBuf scan_tmp;
    DevBuf slice_count;  // uint32_t[1]
    DevBuf root_gate;    // uint32_t[1]  set when some target evaluates the root as a monopole
    DevBuf meta;         // uint32_t[M_WORDS]
    DevBuf rec0, rec1, rec2;   // walk records
    DevBuf recp, bfs_of, recp_par, recp_hist, recp_pref;   // ... in pre-order numbering (sfe_tree.cuh WalkRecP)
    uint32_t* meta_host = nullptr;     // pinned, (SFE_MAX_LEVELS + 2) snapshots of meta
    cudaEvent_t ev_perm[SFE_MAX_LEVELS + 3], ev_side[SFE_SIDE_STREAMS], ev_pre[SFE_SIDE_STREAMS], ev_pre2[SFE_SIDE_STREAMS];
    bool ev_init = false;
};

// ---------------------------------------------------------------------------- SPH cell list (sfe_sph_cells.cu)
struct CellGrid {                 // device-resident descriptor, written by cl_grid_kernel
    uint32_t bmin[3], bmax[3];    // bounding box (order-preserving uint encoding of the floats)
    uint32_t hmin, hmax;          // smoothing-length range (same encoding)
    float origin[3], inv_s, rr;   // cell (0,0,0) corner, 1 / cell size, (R / s)^2 with R = 4 h_max
    uint32_t n[3], ncells, cpr;   // grid dimensions, reach in cells
    uint32_t use_cells, uniform_h;
};
struct SphCells {
    DevBuf desc, count, start, key[2], val[2], hist, rec, tlist;
    DevBuf stat;                  // pairs evaluated (statistics hook)
    DevBuf part, arrive;          // j-split of the pair pass: partial sums of both halves of a group, arrival counters
    bool maybe = false;           // host-side: the cell kernels were enqueued for the current snapshot
};

// ---------------------------------------------------------------------------- resident simulation state (sfe_sim.cu)
constexpr int SFE_SIM_TV_RING = 64;   // steps whose energies may be in flight before sfe_sim_run() synchronises
struct SimState {
    uint32_t N = 0, dof = 0;
    bool loaded = false, scratch_ready = false;
    DevBuf pm4, vel, acc, pot, ken, pen, rad, rho, prs;   // the blocks of sim.mem (simulation.h:17-51); pos + mas live in pm4
    DevBuf stage;                                         // pos3 / mas staging for upload and download
    DevBuf partial, ticket, box_dev, tv_dev;              // reduction scratch, device copy of the root box, {T, V}
    void* box_host = nullptr; void* box_host_dev = nullptr;   // mapped: the root box as the device publishes it
    double* tv_host = nullptr;                            // pinned ring of {T, V}
    uint32_t box_seq = 0;
    float last_box[6] = {0, 0, 0, 0, 0, 0};
    // multi-rank mode (an open peer window with world > 1, sfe_peer.cu): the {x,y,z,m} arrays ping-pong inside the window
    bool peer_mode = false;
    uint32_t i0 = 0, ni = 0;              // target slice of this rank ([0, N) otherwise)
    int cur = 0;                          // which of the two window arrays holds the current positions
    uint64_t pm4_off[2] = {0, 0}, rho_off = 0, acc_off = 0, pot_off = 0;   // window offsets
    uint32_t tv_parity = 0;               // which {T, V} slot set of the control page the next step uses
};

// ---------------------------------------------------------------------------- peer exchange window (sfe_peer.cu)
constexpr uint32_t SFE_PEER_CTRL_BYTES = 4096;   // control page behind the payload: arrival flags, error word, {T, V} slots
struct PeerPtrs { char* base[SFE_PEER_MAX]; uint32_t world, rank; };   // kernel argument: every rank's window
struct PeerCtrl {                          // the control page behind the payload
    uint32_t flag[SFE_PEER_MAX];           // flag[r]: last epoch that rank r has completed (written by rank r)
    uint32_t pad[16];
    double tv[2][SFE_PEER_MAX][4];         // {T.hi, T.lo, V.hi, V.lo} of every rank's slice, ping-pong by step parity
};
static_assert(sizeof(PeerCtrl) <= SFE_PEER_CTRL_BYTES, "control page too small");
struct PeerState {
    bool created = false, open = false;
    uint32_t rank = 0, world = 1;
    uint64_t bytes = 0;                 // payload bytes (the control page follows at this offset)
    void* local = nullptr;              // cudaMalloc'ed window of this rank
    void* base[SFE_PEER_MAX] = {};      // every rank's window as mapped into this process (base[rank] == local)
    uint32_t epoch = 0;                 // exchanges enqueued so far (all ranks enqueue the same sequence)
    uint32_t* err_host = nullptr;       // mapped pinned: set by a wait kernel that timed out
    uint32_t* err_dev = nullptr;
};

// ---------------------------------------------------------------------------- context
struct Ctx {
    int device = -1;
    int sm_count = 0;
    int clock_khz = 0;
    cudaStream_t stream = nullptr;     // engine stream for the host entry points
    cudaStream_t side[SFE_SIDE_STREAMS] = {nullptr, nullptr, nullptr, nullptr};   // overlapped work (moment folds)
    cudaStream_t side2[SFE_SIDE_STREAMS] = {nullptr, nullptr, nullptr, nullptr};  // the folds of the shorter nodes of the top levels, beside the chunked fold
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // device copies of the host arrays
    DevBuf d_pos, d_mas, d_pm4, d_acc, d_pot, d_rad, d_rho, d_prs;
    // PP / SPH scratch
    DevBuf pp_src, pp_partial, pp_flag, sph_src, sph_partial;
    SphCells sph_cells;
    // pinned bounce buffer for unregistered host memory
    void* bounce = nullptr; size_t bounce_cap = 0;
    Tree tree;
    SimState sim;
    PeerState peer;
    // SOLVER_VIS overlay (sfe_overlay.cu): segment list, resident rows of the render buffer, mapped result block
    DevBuf ov_lines, ov_rows;
    // vis_draw() replacement (sfe_vis.cu): integer accumulation buffer, float render buffer, {MVP, mean radius} block
    DevBuf vis_atomic, vis_render, vis_info, vis_pm4;
    void* ov_info = nullptr; void* ov_info_dev = nullptr;
    sfe_timings_t timings;
    uint32_t launches = 0;
    uint64_t total_launches = 0;
};

Ctx* ctx();                 // nullptr before sfe_init
int ensure_init(const char* fn);

// host <-> device staging that copes with pinned and pageable host memory
int h2d(const char* fn, void* dst, const void* src, size_t bytes, cudaStream_t s);
int d2h(const char* fn, void* dst, const void* src, size_t bytes, cudaStream_t s);

// ---------------------------------------------------------------------------- kernels (launchers)
int launch_pack_pm4(const float* pos, const float* mas, float4* pm4, uint32_t n, cudaStream_t s);
int launch_pp(Ctx* c, const float4* pm4, uint32_t N, uint32_t i0, uint32_t ni, float* acc, float* pot,
              float G, float eps2, cudaStream_t s);
// `bounds` is host memory; `bounds_dev` (optional) names the same box in device memory instead - the output of
// launch_root_box_dev() - and spares the host round trip of the six floats
int launch_tree_build(Ctx* c, const float bounds[6], const float4* pm4, uint32_t N, int depth_max,
                      uint32_t nodes_max, cudaStream_t s, const float* bounds_dev = nullptr);
int tree_status(Ctx* c, cudaStream_t s);
int launch_bh(Ctx* c, const float4* pm4, uint32_t N, uint32_t i0, uint32_t ni, float* acc, float* pot,
              float G, float eps2, float theta, uint32_t flavour, cudaStream_t s);
int launch_bh_routed(Ctx* c, const float4* pm4, uint32_t N, uint64_t acc_off, uint64_t pot_off,
                     float G, float eps2, float theta, uint32_t flavour, cudaStream_t s);
// Multi-rank routing of the cell-list pass (sfe_sph_cells.cu): the rank evaluates the targets of a contiguous range of
// SORTED slots (dense in space) and stores rho into every rank's window / adds the pressure force into the owner's acc.
struct SphRoute { char* base[SFE_PEER_MAX]; uint64_t off; uint32_t world, rank; };
int launch_sph_density(Ctx* c, const float4* pm4, const float* rad, uint32_t N, uint32_t i0, uint32_t ni,
                       float* rho, cudaStream_t s, const SphRoute* route = nullptr);
int launch_sph_eos(const float* rho, float* prs, uint32_t n, float eos_k, float eos_n, cudaStream_t s);
int launch_sph_force(Ctx* c, const float4* pm4, const float* rad, const float* rho, const float* prs,
                     uint32_t N, uint32_t i0, uint32_t ni, float cutoff2, float* acc, cudaStream_t s,
                     bool reuse_cells = false, const SphRoute* route = nullptr);
// resident leapfrog (sfe_sim.cu)
int launch_kick_drift(Ctx* c, float4* pm4, float* vel, const float* acc, uint32_t i0, uint32_t ni, float dt, cudaStream_t s);
int launch_root_box(Ctx* c, const float4* pm4, uint32_t N, uint32_t step_number, float bounds[6], cudaStream_t s);
// the same without waiting for the box on the host: *bounds_dev stays valid until the next root box; the host copy can be
// read with root_box_host() once the stream has passed the kernel
int launch_root_box_dev(Ctx* c, const float4* pm4, uint32_t N, uint32_t step_number, const float** bounds_dev, cudaStream_t s);
int root_box_host(Ctx* c, float bounds[6], cudaStream_t s, const char* fn);
int launch_kick_energy(Ctx* c, const float4* pm4, float* vel, const float* acc, const float* pot, float* ken, float* pen,
                       uint32_t i0, uint32_t ni, float dt, int do_kick, double* tv_dev, cudaStream_t s);
int launch_mean_radius(Ctx* c, const float4* pm4, uint32_t N, const double** rbar_dev, cudaStream_t s);
int sim_resident_views(Ctx* c, const float4** pm4, const float** pot, uint32_t* N, const char* fn,
                       uint32_t* i0, uint32_t* ni, uint64_t* vis_off, bool* sliced);
void sim_release(Ctx* c);
// peer exchange (sfe_peer.cu)
PeerPtrs peer_ptrs(const Ctx* c);
int launch_peer_barrier(Ctx* c, cudaStream_t s, const char* fn);          // signal every peer, wait for every peer
int launch_peer_put(Ctx* c, uint64_t offset, uint64_t bytes, cudaStream_t s, const char* fn);
int peer_status(Ctx* c, const char* fn);                                   // raises a timed-out exchange
void peer_release(Ctx* c);
void saver_release();        // sfe_save.cu: drains pending snapshots, stops the writer thread
// SPH neighbour search (sfe_sph_cells.cu): grid + sort for one snapshot, then one gated pass per mode
int cells_stats(Ctx* c, unsigned long long* pairs, cudaStream_t s, const char* fn);
int cells_prepare(Ctx* c, const float4* pm4, const float* rad, uint32_t N, bool force_all_pairs, cudaStream_t s, const char* fn);
int cells_run(Ctx* c, int mode, const float4* pm4, const float* rad, const float* rho, const float* prs, uint32_t N,
              uint32_t i0, uint32_t ni, float cutoff2, float* out, cudaStream_t s, const char* fn, const SphRoute* route = nullptr);

#ifdef __CUDACC__
// ---------------------------------------------------------------- mbarrier / bulk-copy PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}


// bare MUFU.RSQ / MUFU.EX2 (rsqrtf()/exp2f() add denormal fix-up code around the MUFU)
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Compensated accumulator (same recurrence as the reference's Kahan blocks, gravity.c:771-799)
struct Kahan {
    float s, c;
    __device__ __forceinline__ void add(float x) {
        float y = __fsub_rn(x, c);
        float t = __fadd_rn(s, y);
        c = __fsub_rn(__fsub_rn(t, s), y);
        s = t;
    }
};
#endif  // __CUDACC__

inline int check_launch(const char* fn) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(fn, "kernel launch failed: %s", cudaGetErrorString(e)); return SFE_FAILURE; }
    if (ctx()) { ctx()->launches++; ctx()->total_launches++; }
    return SFE_SUCCESS;
}

}  // namespace sfe
