// sfe_tree.cu - Barnes-Hut octree build on the device (replaces tree_build, tree.c:74-318).
//
// The reference inserts particles one at a time (bucket size 8, serial).  The result is a pure function
// of the particle set, which this file reproduces bit-for-bit with a level-by-level MSD construction
// (SURVEY App. B.4):
//
//   level 0   slots = in-box particles in ascending index (stable compaction, tree.c:142-146);
//   level d   every node owns a contiguous slot range [seg.x, seg.y) whose particles are in ascending
//             index.  bodies = range length.  A node with more than 8 bodies is internal: each member
//             picks its octant against mid = 0.5f*(lo+hi) with the reference's tie rule (the first 8
//             members use the strict test of the resident redistribution, tree.c:244-247; later ones
//             the non-strict test of the arriving particle, tree.c:282-285), and the range is split by
//             a STABLE 8-way partition, so each child again holds ascending indices;
//   keys      the octant chosen at level d is digit d (3 bits, x|y<<1|z<<2, at bit 45-3d) of the
//             particle's 48-bit path key; the final slot order is the (key, index)-sorted order;
//   moments   params += {m*x, m*y, m*z, m} is an fp32 left fold in arrival (= ascending index) order in the
//             reference (tree.c:166-169), i.e. a fold over the node's slot range.  It is evaluated
//             exactly that way (one warp per node, products rounded separately, no FMA), because a
//             more accurate reduction changes walk results by 2e-4 rel-RMS at N = 2^20.
//             The folds run on side streams, overlapped with the partitioning of deeper levels; the
//             root's fold (N dependent adds) is only needed by out-of-box particles and is deferred.
//
// Three parts produce that tree:
//   sfe_top.cuh      the upper four levels in ONE pass over the particles (path codes + stable ranks by warp votes, a histogram
//                    table, one-scan node logic, scatter into 32-byte slot records) when the build is large enough;
//   this file        the grid-wide passes, level by level, for cells that stay large below the top levels (and for every level
//                    when the one-pass construction does not apply): rank_kernel (one segmented scan over the whole slot
//                    array per level, warp votes inside a tile, decoupled look-back between tiles), children_kernel,
//                    scatter_kernel - node sizes from 9 to N members go through the same three kernels, node numbering is
//                    breadth-first (children of a level allocated 8 at a time in parent order); and the moment folds
//                    (moments_short / moments_long / the exact chunked fold xl_index + xl_prep + xl_chain) with the
//                    walk records (finalize_kernel);
//   sfe_subtree.cuh  one CTA per node of 9 .. 2048 bodies finishes everything below it in shared memory.
//
// Bound: latency chains, not bandwidth - 231 MB of algorithmic traffic (B_build) in 0.33 ms at N = 2^20.
#include "sfe_internal.cuh"
#include "sfe_tree.cuh"
#include "sfe_scan.cuh"
#include <chrono>
#include <stdlib.h>
#include <vector>

namespace sfe {

// ============================================================================ level 0: in-box compaction
__device__ __forceinline__ bool in_box(const float* b, float4 p) {
    // tree.c:142-146 (inclusive on both faces)
    return !(b[0] > p.x || p.x > b[3] || b[1] > p.y || p.y > b[4] || b[2] > p.z || p.z > b[5]);
}

struct CompactOp {
    const float4* pm4; uint32_t N; TreeDev t;
    __device__ uint32_t n() const { return N; }
    __device__ uint32_t load(uint32_t i) const { return in_box(t.bounds, pm4[i]) ? 1u : 0u; }
    __device__ void store(uint32_t i, uint32_t ex) const {
        const float4 p = pm4[i];
        if (in_box(t.bounds, p)) slot_store(&t.srec[0][ex], p, 0ull, i, 0u);
        else t.perm_out[i - ex] = i;
    }
    __device__ void total(uint32_t M) const {
        // root node (tree.c:113-115) and the per-build counters
        t.meta[M_NUM_INBOX] = M;
        t.meta[M_COUNT] = 1u;
        t.meta[M_STATUS] = 0u;
        t.meta[M_DEPTH] = 0u;
        t.meta[M_FINAL] = 0u;
        t.meta[M_LV_BEGIN + 0] = 0u;
        t.meta[M_LV_END + 0] = 1u;
        WalkRec2 root;                       // root cell (tree.c:113-115)
        for (int k = 0; k < 3; k++) { root.lo[k] = t.bounds[k]; root.hi[k] = t.bounds[3 + k]; }
        root.parent = 0u; root.pad = 0u;
        t.rec2[0] = root;
        t.rec1[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        t.seg[0] = make_uint2(0u, M);
        t.next[0] = NODE_END;
    }
};

// ============================================================================ per level: three kernels
// Level d is processed by three launches that hand over through device memory only:
//   rank_kernel(d)      every slot of an internal level-d node picks its octant (tie rule below) and gets its
//                       stable rank inside (node, octant): ONE pass over the slot array, a segmented scan of
//                       eight counters with decoupled look-back between 2048-slot tiles; also the per-node
//                       octant totals counts[n][8];
//   children_kernel(d)  creates the eight children of every internal level-d node (tree.c:191-274) and, in the
//                       same pass, allocates THEIR children: a look-back scan over the level counts the children
//                       that are internal themselves (more than 8 bodies) and hands each of them a block of
//                       eight node ids of level d+2 in parent order (breadth-first canonical numbering);
//   scatter_kernel(d)   moves every slot to its place in level d+1 (stable 8-way partition) and appends the key digit.
// The root plays the part of children_kernel(-1) (root_kernel).  The only host involvement is "how deep does the
// tree go": children_kernel(d) publishes the number of internal nodes of level d+1 into mapped host memory and the
// host stops enqueuing levels two levels after it reads a zero there (the kernels of an empty level return at once).
constexpr int PT_THREADS = 256;
constexpr int PT_ITEMS = 8;
constexpr int PT_TILE = PT_THREADS * PT_ITEMS;
constexpr uint32_t DESC_WORDS = 20u;   // tile descriptor of the segmented look-back: [0] status, [1] has_head, [2..9] agg, [10..17] incl
constexpr uint32_t ST_NONE = 0u, ST_AGG = 1u, ST_INCL = 2u;
constexpr int CH_THREADS = 256;
constexpr uint32_t ML_MIN = 64u;         // nodes with more members than this take the pipelined fold kernel (moments_long)
constexpr int XL_DEPTH_MAX = 6;          // deeper levels never use the chunked exact fold (their long nodes go through moments_long)

// Octant of a member with the reference's tie rule; mid = 0.5f*(lo+hi) (tree.c:159-163) comes from nmid[].
__device__ __forceinline__ uint32_t pick_octant(float4 mid, float4 p, bool resident) {
    if (resident) return (mid.x < p.x ? 1u : 0u) + (mid.y < p.y ? 2u : 0u) + (mid.z < p.z ? 4u : 0u);     // tree.c:244-247
    return (mid.x <= p.x ? 1u : 0u) + (mid.y <= p.y ? 2u : 0u) + (mid.z <= p.z ? 4u : 0u);                 // tree.c:282-285
}

__device__ __forceinline__ bool level_idle(const TreeDev& t, int d) {   // no internal node at level d (or the build failed)
    return t.meta[M_STATUS] != 0u || t.meta[M_LV_END + d + 1] == t.meta[M_LV_BEGIN + d + 1];
}

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

// ---------------------------------------------------------------------------- rank_kernel
// per-slot byte staged in shared memory between the coalesced load phase and the thread-contiguous scan phase
constexpr uint32_t SB_ACTIVE = 0x08u, SB_HEAD = 0x10u, SB_LAST = 0x20u;

// Block-parallel decoupled look-back of the segmented scan: thread k inspects tile (tile-1-k) of the current window of
// PT_THREADS tiles (warp w its 32 nearest-but-32w); the nearest tile that either already knows its inclusive prefix or
// starts the open range ends the search.  A 2^20-slot level has 512 tiles, so two windows always suffice where a
// one-warp look-back walked up to sixteen windows of 32 one after the other (the 26 us of the old rank_kernel).
// Called by ALL threads of the CTA; every thread returns the same carry.
__device__ __forceinline__ void rank_lookback(const uint32_t* __restrict__ desc_base, uint32_t tile, uint32_t carry[8],
                                              uint32_t (*s_lb)[9] /* [PT_THREADS / 32][9] */) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 0; o < 8; o++) carry[o] = 0u;
    int base = (int)tile - 1;
    for (;;) {
        const int j = base - (int)threadIdx.x;
        const bool valid = j >= 0;
        const uint32_t* dj = desc_base + (size_t)(valid ? j : 0) * DESC_WORDS;
        uint32_t st;
        do { st = valid ? ld_volatile(&dj[0]) : ST_INCL; } while (__any_sync(0xFFFFFFFFu, st == ST_NONE));
        __threadfence();
        const bool stop = !valid || st == ST_INCL || ld_volatile(&dj[1]) != 0u;
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, stop);
        const uint32_t first = ballot ? (uint32_t)__ffs(ballot) - 1u : 32u;
        uint32_t v[8];
#pragma unroll
        for (int o = 0; o < 8; o++) v[o] = 0u;
        if (valid && lane <= first) {
            const uint32_t off = st == ST_INCL ? 10u : 2u;
#pragma unroll
            for (int o = 0; o < 8; o++) v[o] = ld_volatile(&dj[off + o]);
        }
#pragma unroll
        for (int o = 0; o < 8; o++) {
#pragma unroll
            for (int sh = 16; sh >= 1; sh >>= 1) v[o] += __shfl_xor_sync(0xFFFFFFFFu, v[o], sh);
        }
        if (lane == 0u) {
#pragma unroll
            for (int o = 0; o < 8; o++) s_lb[warp][o] = v[o];
            s_lb[warp][8] = ballot ? 1u : 0u;
        }
        __syncthreads();
        bool done = false;
        for (uint32_t w = 0; w < (uint32_t)(PT_THREADS / 32) && !done; w++) {    // nearest warp first
#pragma unroll
            for (int o = 0; o < 8; o++) carry[o] += s_lb[w][o];
            done = s_lb[w][8] != 0u;
        }
        __syncthreads();                                                         // s_lb is reused by the next window
        if (done) return;
        base -= PT_THREADS;
    }
}

// Member masks of a group of 32 slots from four warp votes (active, and the three bits of the octant): the lanes that are
// active in octant (lane & 7) - lanes 0..7 between them hold the eight per-octant masks, and any lane fetches the mask of
// its own slot's octant from lane `octant` with one shuffle (what match_any would return).
__device__ __forceinline__ uint32_t octant_mask(bool act, uint32_t oc, uint32_t lane) {
    const uint32_t ba = __ballot_sync(0xFFFFFFFFu, act);
    const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, (oc & 1u) != 0u), b1 = __ballot_sync(0xFFFFFFFFu, (oc & 2u) != 0u),
                   b2 = __ballot_sync(0xFFFFFFFFu, (oc & 4u) != 0u);
    return ba & ((lane & 1u) ? b0 : ~b0) & ((lane & 2u) ? b1 : ~b1) & ((lane & 4u) ? b2 : ~b2);
}

// One tile = PT_TILE consecutive slots; warp w owns the PT_GROUPS * 32 consecutive slots [w * 256, (w + 1) * 256) of it
// and walks them as PT_GROUPS groups of 32 with lane = slot, so every global access is a coalesced 128-byte line and the
// segmented scan is done with warp votes instead of a scan of eight counters per element:
//   rank inside the group   popc(lanes of my octant & lanes before me & lanes from the last range head on)
//   running counts          lane o (and o + 8, ...) carries the count of octant o of the range that is open at the group
//                           start; a lane fetches the one of its own octant with one shuffle
//   node totals             at the LAST member of a range lanes 0..7 store the eight totals (one 32-byte store)
// Carries between the warps of a tile go through shared memory, between tiles through the decoupled look-back.
// (An earlier version scanned eight 16-bit counters per element through registers, shuffles and shared memory:
// ~230 thread instructions per slot and the longest of the three kernels of a level.)
constexpr int PT_WARPS = PT_THREADS / 32;
constexpr int PT_GROUPS = PT_TILE / PT_THREADS;

__global__ void __launch_bounds__(PT_THREADS, 4) rank_kernel(TreeDev t, int d) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_lb[PT_WARPS][9];
    __shared__ uint32_t s_wtr[PT_WARPS][8];      // per warp: counts of the range that is open at the end of its chunk ...
    __shared__ uint32_t s_whh[PT_WARPS];         // ... counted from the last head inside the chunk, if there is one
    if (level_idle(t, d)) return;
    const uint32_t M = t.meta[M_NUM_INBOX];
    const uint32_t ntiles = (M + PT_TILE - 1) / PT_TILE;
    const SlotRec* __restrict__ rec = t.srec[d];
    uint32_t* __restrict__ desc_base = t.tdesc + (size_t)d * t.tdesc_stride;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, myo = lane & 7u;
    const uint32_t lt = (1u << lane) - 1u, le = lt | (1u << lane);
    for (;;) {
        // tiles are handed out in launch order, so a tile only ever waits for tiles that are already running
        if (threadIdx.x == 0) s_tile = atomicAdd(&t.ticket[2 * d], 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) return;
        const uint32_t w0 = tile * PT_TILE + warp * (PT_GROUPS * 32) + lane;      // this lane's slot in group 0
        // phase 1: octant of every active slot (tie rule) + range head / last-member flags
        uint32_t b[PT_GROUPS], nn[PT_GROUPS], hm[PT_GROUPS], bo[PT_GROUPS];
#pragma unroll
        for (int g = 0; g < PT_GROUPS; g++) {
            const uint32_t s = w0 + 32u * g;
            uint32_t byte = 0u, n = NODE_END;
            if (s < M) {
                const SlotRec r = slot_load(&rec[s]);
                n = r.owner;
                if (n != NODE_END) {                       // slot of an internal node (scatter marks the others)
                    const float4 mid = t.nmid[n];
                    const uint2 sg = t.seg[n];
                    byte = pick_octant(mid, r.pm, s - sg.x < 8u) | SB_ACTIVE | (s == sg.x ? SB_HEAD : 0u) | (s + 1u == sg.y ? SB_LAST : 0u);
                }
            }
            b[g] = byte; nn[g] = n;
        }
        // A tile without a single active slot (most tiles of the deep levels): a range consists of active slots only, so
        // nothing is open at its start or at its end - its inclusive prefix is zero and there is nothing to rank
        {
            uint32_t any = 0u;
#pragma unroll
            for (int g = 0; g < PT_GROUPS; g++) any |= b[g];
            if (__syncthreads_or((int)(any & SB_ACTIVE)) == 0) {
                uint32_t* desc = desc_base + (size_t)tile * DESC_WORDS;
                if (threadIdx.x < 32) {
                    if (threadIdx.x < 8) { desc[2 + threadIdx.x] = 0u; desc[10 + threadIdx.x] = 0u; }
                    if (threadIdx.x == 8) desc[1] = 0u;
                    __threadfence();
                    __syncwarp();
                    if (threadIdx.x == 0) { __threadfence(); atomicExch(&desc[0], ST_INCL); }
                }
                __syncthreads();
                continue;
            }
        }
        // pass A: counts of the range that is still open at the end of the warp's chunk (from its last head on)
        int gl = -1;
#pragma unroll
        for (int g = 0; g < PT_GROUPS; g++) {
            hm[g] = __ballot_sync(0xFFFFFFFFu, (b[g] & SB_HEAD) != 0u);
            if (hm[g]) gl = g;
            bo[g] = octant_mask((b[g] & SB_ACTIVE) != 0u, b[g] & 7u, lane);
        }
        uint32_t tr = 0u;
#pragma unroll
        for (int g = 0; g < PT_GROUPS; g++)
            if (g >= gl) tr += __popc(g == gl ? (bo[g] & (0xFFFFFFFFu << (31 - __clz(hm[g])))) : bo[g]);   // uniform across the warp
        if (lane < 8u) s_wtr[warp][lane] = tr;
        if (lane == 0u) s_whh[warp] = gl >= 0 ? 1u : 0u;
        __syncthreads();
        // what the warps before this one leave open (cw, cut at a head: cf), and the same for the whole tile (tt, tf)
        uint32_t cw = 0u, tt = 0u; bool cf = false, tf = false;
#pragma unroll
        for (uint32_t w = 0; w < (uint32_t)PT_WARPS; w++) {
            const uint32_t v = s_wtr[w][myo]; const bool h = s_whh[w] != 0u;
            if (w < warp) { if (h) { cw = v; cf = true; } else cw += v; }
            if (h) { tt = v; tf = true; } else tt += v;
        }
        // publish the aggregate, fetch the carry of the range that is open at the tile start (decoupled look-back)
        uint32_t* desc = desc_base + (size_t)tile * DESC_WORDS;
        if (threadIdx.x < 8) desc[2 + threadIdx.x] = tt;
        if (threadIdx.x == 8) desc[1] = tf ? 1u : 0u;
        __threadfence();
        __syncthreads();
        uint32_t cm = 0u;                                  // carry into the tile for octant myo
        if (tile > 0u) {                                   // uniform across the CTA
            uint32_t carry[8];
            if (threadIdx.x == 0) atomicExch(&desc[0], ST_AGG);
            rank_lookback(desc_base, tile, carry, s_lb);
#pragma unroll
            for (uint32_t o = 0; o < 8u; o++) if (o == myo) cm = carry[o];
        }
        if (threadIdx.x < 32) {
            if (threadIdx.x < 8) { desc[10 + threadIdx.x] = tt + (tf ? 0u : cm); __threadfence(); }
            __syncwarp();
            if (threadIdx.x == 0) { __threadfence(); atomicExch(&desc[0], ST_INCL); }
        }
        // pass B: ranks, node totals
        uint32_t run = cf ? cw : cw + cm;                  // octant myo of the range open at the start of the warp's chunk
#pragma unroll
        for (int g = 0; g < PT_GROUPS; g++) {
            const uint32_t s = w0 + 32u * g, byte = b[g], oc = byte & 7u;
            const bool act = (byte & SB_ACTIVE) != 0u;
            const uint32_t headmask = hm[g];
            const uint32_t hb = headmask & le;             // heads at or before this lane
            const bool in_open = hb == 0u;                 // my range began before this group
            const uint32_t seg_lo = in_open ? 0u : 31u - (uint32_t)__clz(hb);
            uint32_t r = __popc(__shfl_sync(0xFFFFFFFFu, bo[g], (int)oc) & lt & (0xFFFFFFFFu << seg_lo));   // same octant, same range, before me
            const uint32_t base = __shfl_sync(0xFFFFFFFFu, run, (int)oc);
            if (in_open) r += base;
            if (act) { t.rank[s] = r; t.oct[s] = (uint8_t)oc; }      // scatter_kernel reads them for active slots only
            // last member of a node: the eight totals of its range
            for (uint32_t m = __ballot_sync(0xFFFFFFFFu, act && (byte & SB_LAST) != 0u); m; m &= m - 1u) {
                const uint32_t L = (uint32_t)__ffs(m) - 1u, upto = 0xFFFFFFFFu >> (31u - L);
                const uint32_t hbl = headmask & upto;
                const uint32_t range = hbl ? (upto & (0xFFFFFFFFu << (31 - __clz(hbl)))) : upto;
                const uint32_t val = __popc(bo[g] & range) + (hbl ? 0u : run);
                const uint32_t n = __shfl_sync(0xFFFFFFFFu, nn[g], (int)L);
                if (lane < 8u) t.counts[(size_t)n * 8u + lane] = val;
            }
            // the range that is open at the end of the group
            if (headmask) run = __popc(bo[g] & (0xFFFFFFFFu << (31 - __clz(headmask)))); else run += __popc(bo[g]);
        }
        __syncthreads();   // s_wtr / s_whh / s_lb are reused by the next ticket
    }
}

// a node of 9 .. sub_max bodies is a subtree root: the grid-wide passes stop there, subtree_kernel finishes it
__device__ __forceinline__ bool sub_owns(const TreeDev& t, uint32_t cn) { return cn > 8u && cn <= t.sub_max; }
__device__ __forceinline__ void sub_append(const TreeDev& t, uint32_t node, uint32_t level) {
    t.sub_list[atomicAdd(t.sub_count, 1u)] = make_uint2(node, level);
}

}  // namespace sfe
#include "sfe_subtree.cuh"
namespace sfe {

// ---------------------------------------------------------------------------- children_kernel
// level summary for the host (mapped pinned memory): {build id, internal nodes of the level, status, node count}
__device__ __forceinline__ void publish_summary(const TreeDev& t, int level, uint32_t internal, uint32_t status, uint32_t count) {
    volatile uint32_t* h = t.summary + 4 * level;
    h[1] = internal; h[2] = status; h[3] = count;
    __threadfence_system();
    h[0] = t.build_id;
}

// Bookkeeping shared by root_kernel and the last tile of children_kernel: level `lv` has `internal` internal nodes;
// allocate their children as level lv+1 (tree.c:180-189: node budget first, then depth).
__device__ __forceinline__ void open_level(const TreeDev& t, int lv, uint32_t internal, int depth_max) {
    const uint32_t end = t.meta[M_LV_END + lv];
    uint32_t next_end = end + 8u * internal;
    if (internal > 0u) {
        if (t.nodes_max <= next_end) { t.meta[M_STATUS] = 1u; next_end = end; }
        else if (depth_max <= lv) { t.meta[M_STATUS] = 2u; next_end = end; }
        else t.meta[M_DEPTH] = (uint32_t)lv + 1u;
    }
    t.meta[M_LV_BEGIN + lv + 1] = end;
    t.meta[M_LV_END + lv + 1] = next_end;
    t.meta[M_COUNT] = next_end;
    if (next_end == end) t.meta[M_FINAL] = (uint32_t)lv;   // nothing splits at level lv: its slot order is the final one
    __threadfence();
    publish_summary(t, lv, next_end != end ? internal : 0u, t.meta[M_STATUS], next_end);
}

__global__ void root_kernel(TreeDev t, int depth_max) {
    // runs after the compaction: the root is internal when it holds more than 8 bodies
    const uint32_t M = t.meta[M_NUM_INBOX];
    const bool internal = M > t.sub_max;                 // split by the grid-wide passes
    if (sub_owns(t, M)) sub_append(t, 0u, 0u);           // the whole tree is one subtree
    t.child_base[0] = internal ? 1u : NODE_END;
    const float* b = t.bounds;
    t.nmid[0] = make_float4(__fmul_rn(0.5f, __fadd_rn(b[0], b[3])), __fmul_rn(0.5f, __fadd_rn(b[1], b[4])),
                            __fmul_rn(0.5f, __fadd_rn(b[2], b[5])), 0.f);
    if (M > t.xl_min && !sub_owns(t, M)) { t.xl_count[0] = 1u; t.xl_nodes[0] = 0u; }
    open_level(t, 0, internal ? 1u : 0u, depth_max);
    if (t.meta[M_STATUS] != 0u) t.child_base[0] = NODE_END;
}

__global__ void __launch_bounds__(CH_THREADS) children_kernel(TreeDev t, int d, int depth_max) {
    __shared__ uint32_t sm[33];
    __shared__ uint32_t s_tile, s_prefix;
    __shared__ uint32_t s_cl[CH_THREADS / 32][2];
    if (level_idle(t, d)) {
        // nothing splits at level d: level d+1 does not exist
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (t.meta[M_STATUS] == 0u) { t.meta[M_LV_BEGIN + d + 2] = t.meta[M_LV_END + d + 1]; t.meta[M_LV_END + d + 2] = t.meta[M_LV_END + d + 1]; }
            publish_summary(t, d + 1, 0u, t.meta[M_STATUS], t.meta[M_COUNT]);
        }
        return;
    }
    const uint32_t nb = t.meta[M_LV_BEGIN + d], ne = t.meta[M_LV_END + d];
    const uint32_t ntiles = (ne - nb + CH_THREADS - 1) / CH_THREADS;
    const uint32_t child_level_end = t.meta[M_LV_END + d + 1];
    unsigned long long* __restrict__ cdesc = t.cdesc + (size_t)d * t.cdesc_stride;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&t.ticket[2 * d + 1], 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) return;
        const uint32_t n = nb + tile * CH_THREADS + threadIdx.x;
        const uint32_t base = n < ne ? t.child_base[n] : NODE_END;
        uint32_t cnt[8], ki = 0u;
        if (base != NODE_END) {
            const uint4 c0 = *reinterpret_cast<const uint4*>(&t.counts[(size_t)n * 8u]);
            const uint4 c1 = *reinterpret_cast<const uint4*>(&t.counts[(size_t)n * 8u + 4u]);
            cnt[0] = c0.x; cnt[1] = c0.y; cnt[2] = c0.z; cnt[3] = c0.w; cnt[4] = c1.x; cnt[5] = c1.y; cnt[6] = c1.z; cnt[7] = c1.w;
#pragma unroll
            for (int o = 0; o < 8; o++) ki += cnt[o] > t.sub_max ? 1u : 0u;
        }
        uint32_t tile_total;
        const uint32_t ex = block_excl_scan(ki, &tile_total, sm);
        // block-parallel decoupled look-back over the tiles of this level (status in the high word, value in the low):
        // thread k inspects tile (tile-1-k); a deep level has more than a thousand tiles of CH_THREADS nodes, which a
        // one-warp look-back walked in windows of 32 one after the other
        {
            const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
            uint32_t prefix = 0u;
            if (tile > 0u) {                                       // uniform across the CTA
                if (threadIdx.x == 0) atomicExch(&cdesc[tile], ((unsigned long long)ST_AGG << 32) | tile_total);
                int jb = (int)tile - 1;
                for (;;) {
                    const int j = jb - (int)threadIdx.x;
                    const bool valid = j >= 0;
                    unsigned long long v;
                    do { v = valid ? *reinterpret_cast<volatile unsigned long long*>(&cdesc[j]) : ((unsigned long long)ST_INCL << 32); }
                    while (__any_sync(0xFFFFFFFFu, (uint32_t)(v >> 32) == ST_NONE));
                    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, (uint32_t)(v >> 32) == ST_INCL);
                    const uint32_t first = ballot ? (uint32_t)__ffs(ballot) - 1u : 32u;
                    uint32_t x = lane <= first ? (uint32_t)v : 0u;
#pragma unroll
                    for (int sh = 16; sh >= 1; sh >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, sh);
                    if (lane == 0u) { s_cl[warp][0] = x; s_cl[warp][1] = ballot ? 1u : 0u; }
                    __syncthreads();
                    bool done = false;
                    for (uint32_t w = 0; w < (uint32_t)(CH_THREADS / 32) && !done; w++) { prefix += s_cl[w][0]; done = s_cl[w][1] != 0u; }
                    __syncthreads();
                    if (done) break;
                    jb -= CH_THREADS;
                }
            }
            if (threadIdx.x == 0) {
                atomicExch(&cdesc[tile], ((unsigned long long)ST_INCL << 32) | (prefix + tile_total));
                s_prefix = prefix;
                if (tile == ntiles - 1u) open_level(t, d + 1, prefix + tile_total, depth_max);   // level d+1 has this many internal nodes
            }
        }
        __syncthreads();
        if (base != NODE_END) {
            const float4 b0 = *reinterpret_cast<const float4*>(&t.rec2[n]);          // lo.x lo.y lo.z hi.x
            const float4 b1 = *(reinterpret_cast<const float4*>(&t.rec2[n]) + 1);    // hi.y hi.z parent -
            const float lo[3] = {b0.x, b0.y, b0.z}, hi[3] = {b0.w, b1.x, b1.y};
            float mid[3];
#pragma unroll
            for (int k = 0; k < 3; k++) mid[k] = __fmul_rn(0.5f, __fadd_rn(lo[k], hi[k]));   // tree.c:159-163
            uint32_t start = t.seg[n].x;
            const uint32_t nxt = t.next[n];
            uint32_t r = s_prefix + ex;                  // rank of this node's first internal child within level d+1
#pragma unroll
            for (uint32_t o = 0; o < 8; o++) {
                const uint32_t c = base + o, cn = cnt[o];
                float clo[3], chi[3];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    clo[k] = (o >> k) & 1u ? mid[k] : lo[k];          // tree.c:219-226
                    chi[k] = (o >> k) & 1u ? hi[k] : mid[k];
                }
                float4* cell = reinterpret_cast<float4*>(&t.rec2[c]);
                cell[0] = make_float4(clo[0], clo[1], clo[2], chi[0]);
                cell[1] = make_float4(chi[1], chi[2], __uint_as_float(n), 0.f);
                t.rec1[c] = make_float4(0.f, 0.f, 0.f, 0.f);          // moments of empty nodes stay zero (tree.c:302-310)
                if (cn > t.sub_max) {
                    t.child_base[c] = child_level_end + 8u * r; r++;
                    t.nmid[c] = make_float4(__fmul_rn(0.5f, __fadd_rn(clo[0], chi[0])), __fmul_rn(0.5f, __fadd_rn(clo[1], chi[1])),
                                            __fmul_rn(0.5f, __fadd_rn(clo[2], chi[2])), 0.f);
                } else {
                    t.child_base[c] = NODE_END;
                }
                t.seg[c] = make_uint2(start, start + cn);
                if (sub_owns(t, cn)) sub_append(t, c, (uint32_t)d + 1u);   // everything below it is subtree_kernel's
                if (!sub_owns(t, cn) && d + 1 <= XL_DEPTH_MAX && cn > t.xl_min) {   // very long fold: chunked exact evaluation (moments_xl)
                    const uint32_t k = atomicAdd(&t.xl_count[d + 1], 1u);
                    if (k < t.xl_max) t.xl_nodes[(size_t)(d + 1) * t.xl_max + k] = c;
                } else if (cn > ML_MIN) {
                    t.ml_count[d + 1] = 1u;                       // level d+1 has work for moments_long_kernel (any writer will do)
                }
                // depth-first successor when c is not opened: the next NON-EMPTY sibling, else the parent's successor.
                // The reference only steps back to the parent at an empty node (gravity.c:324-328), so leaving empty
                // nodes out of the walk links changes neither the visit order nor the sums.
                uint32_t succ = nxt;
#pragma unroll
                for (uint32_t o2 = 7u; o2 > o; o2--) if (cnt[o2] > 0u) succ = base + o2;
                t.next[c] = succ;
                start += cn;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------- scatter_kernel
__global__ void __launch_bounds__(256) scatter_kernel(TreeDev t, int d) {
    if (level_idle(t, d)) return;
    const uint32_t M = t.meta[M_NUM_INBOX];
    const SlotRec* __restrict__ rec = t.srec[d];
    SlotRec* __restrict__ rec_next = t.srec[d + 1];
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < M; s += gridDim.x * blockDim.x) {
        const SlotRec r = slot_load(&rec[s]);
        const uint32_t n = r.owner;
        uint32_t dest = s, owner = NODE_END;
        unsigned long long key = r.key;
        if (n != NODE_END) {
            const uint32_t o = t.oct[s];
            const uint32_t c = t.child_base[n] + o;
            const uint2 sg = t.seg[c];
            dest = sg.x + t.rank[s];
            owner = (sg.y - sg.x) > t.sub_max ? c : NODE_END;   // only slots of nodes the grid-wide passes split stay active
            key |= (unsigned long long)o << (45 - 3 * d);   // digit d of the 48-bit path key
        }
        slot_store(&rec_next[dest], r.pm, key, r.idx, owner);
    }
}

// path keys by particle index (test hook): in-box particles from their final slot, the others get all ones
__global__ void keys_gather_kernel(TreeDev t, uint32_t N) {
    const uint32_t M = t.meta[M_NUM_INBOX], fin = t.meta[M_FINAL];
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < N; s += gridDim.x * blockDim.x) {
        if (s < M) { const SlotRec r = slot_load(&t.srec[fin][s]); t.keys[r.idx] = r.key; }
        else t.keys[t.perm_out[s - M]] = ~0ull;
    }
}

__global__ void perm_gather_kernel(TreeDev t, uint32_t* __restrict__ out) {
    const uint32_t M = t.meta[M_NUM_INBOX], fin = t.meta[M_FINAL];
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < M; s += gridDim.x * blockDim.x) out[s] = t.srec[fin][s].idx;
}

// ============================================================================ moments (tree.c:166-169, :302-310)
// params += {m*x, m*y, m*z, m} is an fp32 LEFT FOLD in arrival (= slot) order, so every node is a serial
// dependent-add chain (FADD latency: 4 cycles per member) that no reduction tree may replace.
// One lane per (node, component): a warp carries 32 chains (8 nodes) at once and the 8 depth-1 cells - the
// longest chains of a build, N/8 members each - share ONE warp.  The members of a node are contiguous in
// spm4[d].  Two kernels per level:
//   moments_long_kernel   nodes with more than ML_MIN members.  Warp-specialised CTA: eight PRODUCER warps
//                         stream the raw float4 records into a shared-memory ring with cp.async (LDGSTS),
//                         ML_R batches deep so HBM/L2 latency is hidden, round the products m*x, m*y, m*z
//                         (no FMA) and park them component-major; the CONSUMER warp does nothing but
//                         LDS.128 + 4 dependent FADDs per 4 members, i.e. the chain runs at the FADD latency.
//   moments_short_kernel  the (many) small nodes: one lane per (node, component) straight from global memory.
constexpr int ML_B = 128;                // members per node per batch
constexpr int ML_R = 4;                  // raw ring depth in batches
constexpr int ML_BP = ML_B + 4;          // padded row of the product buffer: conflict-free LDS.128 across 32 rows
constexpr int ML_THREADS = 288;          // warp 0 consumer, warps 1-8 producers (one node each)
constexpr size_t ML_SMEM = (size_t)ML_R * 8 * ML_B * 16 + (size_t)2 * 32 * ML_BP * 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// which kernel folds a node with cn members at level d
// (a subtree root is folded here like any other node of its level - only the TREE root, when the whole tree is one subtree,
// is folded by subtree_kernel itself, because the deferred root fold would find its slots already reordered)
__device__ __forceinline__ bool sub_folds(const TreeDev& t, int d, uint32_t cn) { return d == 0 && sub_owns(t, cn); }
__device__ __forceinline__ bool xl_owns(const TreeDev& t, int d, uint32_t cn) { return d <= XL_DEPTH_MAX && cn > t.xl_min && !sub_owns(t, cn); }   // (its chains run beside subtree_kernel)
__device__ __forceinline__ bool ml_owns(const TreeDev& t, int d, uint32_t cn) { return cn > ML_MIN && !xl_owns(t, d, cn) && !sub_folds(t, d, cn); }

// (all fold kernels take a range of levels d0..d1: the top-level path folds levels 1..TOP_L with one launch of each kernel -
// with four launches per level the host, not the device, was what the folds waited for)
__global__ void __launch_bounds__(ML_THREADS) moments_long_kernel(TreeDev t, int d0, int d1,
                                                                  uint32_t* __restrict__ gate /* {needed, done} or null */) {
    extern __shared__ __align__(16) unsigned char ml_smem[];
    float4 (*raw)[8][ML_B] = reinterpret_cast<float4 (*)[8][ML_B]>(ml_smem);                                  // [ML_R][8][ML_B]
    float (*prod)[32][ML_BP] = reinterpret_cast<float (*)[32][ML_BP]>(ml_smem + (size_t)ML_R * 8 * ML_B * 16);  // [2][32][ML_BP]
    if (gate && (gate[0] == 0u || gate[1] != 0u)) return;
    // children_kernel(d-1) notes whether level d has a node for this kernel at all: deep levels have hundreds of thousands of
    // nodes and none that long, and finding that out group by group costs a 100 us of dependent loads (level 7 at C2)
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int d = d0; d <= d1; d++) {
    if (d > 0 && t.ml_count[d] == 0u) continue;
    const uint32_t b = t.meta[M_LV_BEGIN + d], e = t.meta[M_LV_END + d];
    const SlotRec* __restrict__ rec = t.srec[d];
    // (the groups of a level are dealt out from a different CTA on every level, so that the few groups of the top levels
    // do not all land on the same CTAs)
    for (uint32_t base = b + ((blockIdx.x + (uint32_t)d * 37u) % gridDim.x) * 8u; base < e; base += gridDim.x * 8u) {
        // every thread looks at the 8 segments of the group (uniform loads) for the trip count
        uint32_t maxcnt = 0u;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint2 sg = base + i < e ? t.seg[base + i] : make_uint2(0u, 0u);
            const uint32_t cn = sg.y - sg.x;
            maxcnt = max(maxcnt, ml_owns(t, d, cn) ? cn : 0u);
        }
        if (maxcnt == 0u) continue;
        const uint32_t nbatch = (maxcnt + ML_B - 1) / ML_B;
        if (warp == 0u) {
            // ---------------------------------------------------------------- consumer: 32 chains
            float sum = 0.f;
            for (uint32_t k = 0; k < nbatch; k++) {
                __syncthreads();                                   // prod[k & 1] is complete
                const float4* row = reinterpret_cast<const float4*>(&prod[k & 1u][lane][0]);
#pragma unroll
                for (int u = 0; u < ML_B / 4; u++) {
                    const float4 v = row[u];                       // pads are +0: s + 0 == s
                    sum = __fadd_rn(sum, v.x); sum = __fadd_rn(sum, v.y);   // tree.c:166-169
                    sum = __fadd_rn(sum, v.z); sum = __fadd_rn(sum, v.w);
                }
            }
            __syncthreads();                                       // group done: buffers may be reused
            const uint32_t q = lane >> 2, c = lane & 3u;
            const uint2 msg = base + q < e ? t.seg[base + q] : make_uint2(0u, 0u);
            const uint32_t mycount = ml_owns(t, d, msg.y - msg.x) ? (msg.y - msg.x) : 0u;
            const float mass = __shfl_sync(0xFFFFFFFFu, sum, lane | 3u);
            if (mycount) {
                const float inv = 0.f < mass ? __fdiv_rn(1.0f, mass) : 1.0f;   // tree.c:303-304
                reinterpret_cast<float*>(&t.rec1[base + q])[c] = c < 3u ? __fmul_rn(sum, inv) : sum;
            }
        } else {
            // ---------------------------------------------------------------- producers: one node per warp
            const uint32_t node = warp - 1u;
            const uint2 sg = base + node < e ? t.seg[base + node] : make_uint2(0u, 0u);
            const uint32_t cnt = ml_owns(t, d, sg.y - sg.x) ? (sg.y - sg.x) : 0u;
            const SlotRec* __restrict__ src = rec + sg.x;
            auto issue = [&](uint32_t k) {       // raw batch k -> ring slot k % ML_R
                const uint32_t j0 = k * ML_B;
                if (j0 < cnt) {
                    float4* dst = &raw[k % ML_R][node][0];
#pragma unroll
                    for (int h = 0; h < ML_B / 32; h++) {
                        const uint32_t j = lane + 32u * h;
                        if (j0 + j < cnt) cp_async16(dst + j, &src[j0 + j].pm);
                    }
                }
                cp_async_commit();               // one (possibly empty) group per batch index keeps the wait arithmetic uniform
            };
            auto transpose = [&](uint32_t k) {   // raw batch k -> prod[k & 1], products rounded one by one
                const uint32_t j0 = k * ML_B;
                const float4* rw = &raw[k % ML_R][node][0];
                float* pr = &prod[k & 1u][node * 4u][0];
#pragma unroll
                for (int h = 0; h < ML_B / 32; h++) {
                    const uint32_t j = lane + 32u * h;
                    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j0 + j < cnt) r = rw[j];
                    pr[j] = __fmul_rn(r.w, r.x);
                    pr[ML_BP + j] = __fmul_rn(r.w, r.y);
                    pr[2 * ML_BP + j] = __fmul_rn(r.w, r.z);
                    pr[3 * ML_BP + j] = r.w;
                }
            };
#pragma unroll 1
            for (uint32_t k = 0; k < (uint32_t)ML_R; k++) issue(k);
            cp_async_wait<ML_R - 1>();
            __syncwarp();
            transpose(0u);
            __syncwarp();
            issue((uint32_t)ML_R);
#pragma unroll 1
            for (uint32_t k = 0; k < nbatch; k++) {
                __syncthreads();                                   // hands prod[k & 1] to the consumer
                if (k + 1u < nbatch) {
                    cp_async_wait<ML_R - 1>();                     // raw batch k+1 has landed
                    __syncwarp();
                    transpose(k + 1u);
                    __syncwarp();
                    issue(k + 1u + ML_R);
                }
            }
            cp_async_wait<0>();
            __syncthreads();
        }
    }
    }
}

__global__ void __launch_bounds__(256) moments_short_kernel(TreeDev t, int d0, int d1, uint32_t* __restrict__ gate) {
    if (gate && (gate[0] == 0u || gate[1] != 0u)) return;
    const uint32_t c = threadIdx.x & 3u;
    for (int d = d0; d <= d1; d++) {
    const uint32_t b = t.meta[M_LV_BEGIN + d], e = t.meta[M_LV_END + d];
    const SlotRec* __restrict__ rec = t.srec[d];
    const uint32_t quads = (e - b);
    for (uint32_t qd = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; qd < ((quads + 7u) & ~7u); qd += (gridDim.x * blockDim.x) >> 2) {
        const uint32_t n = b + qd;
        const bool live = n < e;
        const uint2 sg = live ? t.seg[n] : make_uint2(0u, 0u);
        uint32_t cnt = sg.y - sg.x;
        if (cnt > ML_MIN || sub_folds(t, d, cnt)) cnt = 0u;       // the pipelined kernel owns the long nodes, subtree_kernel the tree root it owns
        const float* __restrict__ src = reinterpret_cast<const float*>(rec + sg.x);   // {x, y, z, m} lead every 32-byte record
        float sum = 0.f;
#pragma unroll 4
        for (uint32_t j = 0; j < cnt; j++) {
            const float m = __ldg(src + 8u * j + 3u);
            const float p = __fmul_rn(m, __ldg(src + 8u * j + c));  // tree.c:166-168: product rounded, then added
            sum = __fadd_rn(sum, c == 3u ? m : p);                  // tree.c:169
        }
        const float mass = __shfl_sync(0xFFFFFFFFu, sum, (threadIdx.x & 31u) | 3u);
        if (cnt) {
            const float inv = 0.f < mass ? __fdiv_rn(1.0f, mass) : 1.0f;   // tree.c:303-304
            reinterpret_cast<float*>(&t.rec1[n])[c] = c < 3u ? __fmul_rn(sum, inv) : sum;
        }
    }
    }
}

// ---------------------------------------------------------------------------- exact chunked fold of very long nodes
// The left fold s <- RN(s + p_j) is serial, but while s stays inside one binade [2^23, 2^24) * u every addition is an
// integer step on k = s / u:  k <- RNE(k + p_j / u) = k + floor(t_j) + [frac(t_j) > 1/2] (+1 on an exact tie when
// k + floor(t_j) is odd).  As a function of k such a step only depends on the PARITY of k, steps of that kind are
// closed under composition, and composition is associative - so a whole chunk of XL_CHUNK members collapses, in
// parallel, into a map k -> k + a[k & 1], together with the extremes of the running value inside the chunk.  (All of
// these are below 2^25 in magnitude whenever the map can apply at all, so a map is eight 32-bit words - one sector.)
//   xl_index   one CTA: the level's long nodes with their chunk counts, as a flat list of chunks;
//   xl_prep    one warp per chunk, ONE pass over the members: double sums of the rounded products, their prefix over
//              the node's chunks by decoupled look-back (chunks are handed out in order by a ticket) - the approximate
//              value of the fold at both ends of the chunk, good enough to GUESS its binade, i.e. u - and, with the
//              members still in registers, the parity map and the extremes for that u (a lane composes its eight
//              consecutive members, an ordered warp reduction composes the lanes);
//   xl_chain   the only serial part, one warp per (node, component): the maps of 32 consecutive chunks are composed
//              with a warp scan (lane i: chunks 0..i of the window, as long as they share u), every lane tests its
//              prefix against the actual s - a multiple of u strictly inside the binade, with the extremes keeping
//              every intermediate value inside it - and the longest valid prefix is applied at once, so a run of chunks
//              inside one binade costs one step per 32 chunks.  A chunk whose map does not apply (binade crossings,
//              the first chunks, wrong guesses) is folded member by member.  The reference's bit pattern either way.
constexpr int XL_CHUNK = 256;
constexpr int XL_PER_LANE = XL_CHUNK / 32;
constexpr int XL_LIM = 1 << 26;                 // bound on every entry of a usable map
struct __align__(32) XlMap { int a0, a1, mn0, mx0, mn1, mx1, e2, usable; };
struct __align__(32) XlNode { uint32_t n, first, cnt, nchunks, slot0, choff, pad0, pad1; };   // choff: chunks of the nodes listed before

__device__ __forceinline__ XlMap xl_identity(int e2, int usable) {
    XlMap m;
    m.a0 = m.a1 = 0; m.mn0 = m.mn1 = XL_LIM; m.mx0 = m.mx1 = -XL_LIM; m.e2 = e2; m.usable = usable;
    return m;
}
__device__ __forceinline__ XlMap xl_combine(const XlMap& A, const XlMap& B) {   // A first, then B
    XlMap C;
    const bool p0 = (A.a0 & 1) != 0;              // parity after A when entering even
    const bool p1 = ((1 + A.a1) & 1) != 0;        // parity after A when entering odd
    C.a0 = A.a0 + (p0 ? B.a1 : B.a0);
    C.a1 = A.a1 + (p1 ? B.a1 : B.a0);
    C.mn0 = min(A.mn0, A.a0 + (p0 ? B.mn1 : B.mn0));
    C.mx0 = max(A.mx0, A.a0 + (p0 ? B.mx1 : B.mx0));
    C.mn1 = min(A.mn1, A.a1 + (p1 ? B.mn1 : B.mn0));
    C.mx1 = max(A.mx1, A.a1 + (p1 ? B.mx1 : B.mx0));
    C.e2 = A.e2;
    const int big = max(max(abs(C.a0), abs(C.a1)), max(max(abs(C.mn0), abs(C.mx0)), max(abs(C.mn1), abs(C.mx1))));
    C.usable = (A.usable != 0 && B.usable != 0 && big <= XL_LIM) ? 1 : 0;
    if (!C.usable) { C.a0 = C.a1 = 0; C.mn0 = C.mn1 = 0; C.mx0 = C.mx1 = 0; }   // cannot apply any more; keeps later sums small
    return C;
}
__device__ __forceinline__ XlMap xl_shfl_down(const XlMap& m, int off) {
    XlMap o;
    o.a0 = __shfl_down_sync(0xFFFFFFFFu, m.a0, off); o.a1 = __shfl_down_sync(0xFFFFFFFFu, m.a1, off);
    o.mn0 = __shfl_down_sync(0xFFFFFFFFu, m.mn0, off); o.mx0 = __shfl_down_sync(0xFFFFFFFFu, m.mx0, off);
    o.mn1 = __shfl_down_sync(0xFFFFFFFFu, m.mn1, off); o.mx1 = __shfl_down_sync(0xFFFFFFFFu, m.mx1, off);
    o.e2 = m.e2; o.usable = __shfl_down_sync(0xFFFFFFFFu, m.usable, off);
    return o;
}
__device__ __forceinline__ XlMap xl_shfl_up(const XlMap& m, int off) {
    XlMap o;
    o.a0 = __shfl_up_sync(0xFFFFFFFFu, m.a0, off); o.a1 = __shfl_up_sync(0xFFFFFFFFu, m.a1, off);
    o.mn0 = __shfl_up_sync(0xFFFFFFFFu, m.mn0, off); o.mx0 = __shfl_up_sync(0xFFFFFFFFu, m.mx0, off);
    o.mn1 = __shfl_up_sync(0xFFFFFFFFu, m.mn1, off); o.mx1 = __shfl_up_sync(0xFFFFFFFFu, m.mx1, off);
    o.e2 = m.e2; o.usable = __shfl_up_sync(0xFFFFFFFFu, m.usable, off);
    return o;
}
__device__ __forceinline__ float xl_term(float4 p, uint32_t q) {   // the q-th summand of tree.c:166-169
    return q == 0u ? __fmul_rn(p.w, p.x) : q == 1u ? __fmul_rn(p.w, p.y) : q == 2u ? __fmul_rn(p.w, p.z) : p.w;
}

// the level's long nodes as a flat list of chunks: descriptors + running chunk count, [xl_max] holds the total
__global__ void __launch_bounds__(1024) xl_index_kernel(TreeDev t, int d0, uint32_t* __restrict__ gate) {
    __shared__ uint32_t sm[33];
    __shared__ uint32_t carry;
    if (gate && (gate[0] == 0u || gate[1] != 0u)) return;
    const int d = d0 + (int)blockIdx.x;                 // one CTA per level
    if (d > XL_DEPTH_MAX) return;
    const uint32_t area = (uint32_t)(d % SFE_SIDE_STREAMS);
    XlNode* __restrict__ list = reinterpret_cast<XlNode*>(t.xl_list) + (size_t)area * (t.xl_max + 1u);
    const uint32_t nxl = min(t.xl_count[d], t.xl_max);
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < nxl; base += blockDim.x) {
        const uint32_t k = base + threadIdx.x;
        XlNode nd;
        nd.n = 0u; nd.first = 0u; nd.cnt = 0u; nd.nchunks = 0u; nd.slot0 = 0u; nd.choff = 0u; nd.pad0 = 0u; nd.pad1 = 0u;
        if (k < nxl) {
            nd.n = t.xl_nodes[(size_t)d * t.xl_max + k];
            const uint2 sg = t.seg[nd.n];
            nd.first = sg.x; nd.cnt = sg.y - sg.x;
            nd.nchunks = (nd.cnt + XL_CHUNK - 1) / XL_CHUNK;
            // disjoint chunk-slot ranges, independent of the (arbitrary) list order: a long node has more than xl_min
            // members, so the next one starts at least one xl_min block later
            nd.slot0 = sg.x / XL_CHUNK + 3u * (sg.x / t.xl_min);
        }
        uint32_t tot;
        const uint32_t ex = block_excl_scan(nd.nchunks, &tot, sm);
        nd.choff = carry + ex;
        if (k < nxl) list[k] = nd;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        XlNode end;
        end.n = nxl; end.first = 0u; end.cnt = 0u; end.nchunks = 0u; end.slot0 = 0u; end.choff = carry; end.pad0 = 0u; end.pad1 = 0u;
        list[t.xl_max] = end;                       // {number of long nodes, -, -, -, -, total chunks}
    }
}

constexpr uint32_t XLF_AGG = 1u, XLF_INCL = 2u;   // chunk flag = epoch << 2 | state; other epochs read as "nothing yet"
constexpr int XP_WARPS = 8;
constexpr uint32_t XP_CHOFF = 512u;
constexpr int XP_ROW = 33;                        // staged members: one float4 of padding per 32 (conflict-free strided reads)

__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// One warp per chunk.  The 256 members are staged in shared memory with coalesced loads; then lane 8g + k (g = component
// 0..3, k = 0..7) owns the 32 consecutive members 32k .. 32k+31 of component g: their double sum (for the prefix), and - once
// the look-back has delivered the approximate fold value at the chunk's start - their parity map, composed in order over
// the eight lanes of the group.  Four components in four lane groups: a quarter of the shuffles and no members in registers.
__global__ void __launch_bounds__(XP_WARPS * 32) xl_prep_kernel(TreeDev t, int d0, int d1, uint32_t epoch0, uint32_t* __restrict__ gate) {
    __shared__ float4 stage[XP_WARPS][8 * XP_ROW];
    __shared__ uint32_t s_choff[XP_CHOFF];                     // chunk offsets of the level's long nodes (the first XP_CHOFF of them)
    if (gate && (gate[0] == 0u || gate[1] != 0u)) return;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, q = lane >> 3, sub = lane & 7u;
    const uint32_t qv = q < 3u ? q : 3u;                         // float of {x,y,z,m} that goes into this lane group's product
    float4* __restrict__ st = stage[warp];
    for (int d = d0; d <= d1 && d <= XL_DEPTH_MAX; d++) {      // a warp that finds one level's chunks dealt out moves on to the next
    const uint32_t area = (uint32_t)(d % SFE_SIDE_STREAMS), epoch = (epoch0 + (uint32_t)(d - d0)) & 0x3FFFFFFFu;
    const XlNode* __restrict__ list = reinterpret_cast<const XlNode*>(t.xl_list) + (size_t)area * (t.xl_max + 1u);
    const uint32_t nxl = list[t.xl_max].n, total = list[t.xl_max].choff;
    double* __restrict__ sums_all = t.xl_sum + (size_t)area * t.xl_slots * 8u;           // per chunk slot: aggregate[4], inclusive prefix[4]
    uint32_t* __restrict__ flags = t.xl_flag + (size_t)area * t.xl_slots;
    XlMap* __restrict__ maps_all = reinterpret_cast<XlMap*>(t.xl_map) + (size_t)area * t.xl_slots * 4u;
    uint32_t* ticket = &t.xl_ticket[d];
    __syncthreads();                                             // (every warp of the CTA passes here once per level)
    for (uint32_t k = threadIdx.x; k < min(nxl, XP_CHOFF); k += blockDim.x) s_choff[k] = list[k].choff;
    __syncthreads();
    for (;;) {
        // chunks are handed out in list order, so a chunk only ever waits for chunks that are already being worked on
        uint32_t w = 0u;
        if (lane == 0u) w = atomicAdd(ticket, 1u);
        w = __shfl_sync(0xFFFFFFFFu, w, 0);
        if (w >= total) break;
        uint32_t lo = 0u, hi = nxl;                 // the node whose chunks include w: last k with choff[k] <= w
        while (hi - lo > 1u) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((mid < XP_CHOFF ? s_choff[mid] : list[mid].choff) <= w) lo = mid; else hi = mid;
        }
        const XlNode nd = list[lo];
        const uint32_t c = w - nd.choff;
        const SlotRec* __restrict__ src = t.srec[d] + nd.first;
        const size_t slot = (size_t)nd.slot0 + c;
        const uint32_t left = nd.cnt - c * XL_CHUNK;                 // members of this chunk (>= 1)
        __syncwarp();                                                // the previous chunk's staged members are no longer read
#pragma unroll
        for (uint32_t i = 0; i < (uint32_t)XL_PER_LANE; i++) {       // member 32 i + lane -> row i (pads are {0,0,0,0}: +0 summands)
            const uint32_t j = i * 32u + lane;
            st[i * XP_ROW + lane] = j < left ? src[c * XL_CHUNK + j].pm : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        // this lane's 32 summands of component q: members 32 sub .. 32 sub + 31, i.e. row `sub`
        // (scalar reads at a lane-dependent offset: a select on the component would be a branch per lane group)
        const float* __restrict__ mine = reinterpret_cast<const float*>(st + sub * XP_ROW);
        const uint32_t have = left > sub * 32u ? min(32u, left - sub * 32u) : 0u;
        double a = 0.0;
#pragma unroll 8
        for (uint32_t i = 0; i < 32u; i++) {
            const float m = mine[4u * i + 3u], v = mine[4u * i + qv];
            a += (double)(q == 3u ? m : __fmul_rn(m, v));
        }
        a += __shfl_xor_sync(0xFFFFFFFFu, a, 1); a += __shfl_xor_sync(0xFFFFFFFFu, a, 2); a += __shfl_xor_sync(0xFFFFFFFFu, a, 4);
        // publish the aggregate (the inclusive prefix at once for a node's first chunk), then look back
        if (sub == 0u) { sums_all[slot * 8u + q] = a; if (c == 0u) sums_all[slot * 8u + 4u + q] = a; }
        __syncwarp();
        if (lane == 0u) st_release_u32(&flags[slot], (epoch << 2) | (c == 0u ? XLF_INCL : XLF_AGG));
        double pre = 0.0;                                            // component q: approximate fold value at the chunk's start
        if (c > 0u) {
            int jb = (int)c - 1;                    // nearest predecessor of the current window
            for (;;) {
                const int j = jb - (int)lane;
                const bool valid = j >= 0;
                uint32_t stt;
                do {
                    const uint32_t f = valid ? ld_acquire_u32(&flags[(size_t)nd.slot0 + (uint32_t)j]) : ((epoch << 2) | XLF_INCL);
                    stt = (f >> 2) == epoch ? (f & 3u) : 0u;
                    if (stt == 0u) __nanosleep(40);                  // (a spinning warp takes issue slots from the ones it waits for)
                } while (__any_sync(0xFFFFFFFFu, stt == 0u));
                const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, !valid || stt == XLF_INCL);
                const uint32_t first = (uint32_t)__ffs((int)ballot) - 1u;          // nearest chunk that knows its inclusive prefix (or the node's start)
                double v[4] = {0.0, 0.0, 0.0, 0.0};
                if (valid && lane <= first) {
                    const double* sp = sums_all + ((size_t)nd.slot0 + (uint32_t)j) * 8u + (stt == XLF_INCL ? 4u : 0u);
#pragma unroll
                    for (int k = 0; k < 4; k++) v[k] = __ldcg(sp + k);
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int sh = 16; sh >= 1; sh >>= 1) v[k] += __shfl_xor_sync(0xFFFFFFFFu, v[k], sh);
                pre += q == 0u ? v[0] : q == 1u ? v[1] : q == 2u ? v[2] : v[3];
                if (ballot) break;
                jb -= 32;
            }
            if (sub == 0u) sums_all[slot * 8u + 4u + q] = pre + a;
            __syncwarp();
            if (lane == 0u) st_release_u32(&flags[slot], (epoch << 2) | XLF_INCL);
        }
        // maps: binade guess from the approximate fold value at both ends of the chunk (uniform inside a lane group)
        const float fs = (float)pre, fe = (float)(pre + a);
        const uint32_t es = (__float_as_uint(fs) >> 23) & 0xFFu, ee = (__float_as_uint(fe) >> 23) & 0xFFu;
        const bool usable = es == ee && es > 40u && es < 250u && ((__float_as_uint(fs) ^ __float_as_uint(fe)) >> 31) == 0u;
        XlMap m = xl_identity((int)es - 127 - 23, usable ? 1 : 0);
        {
            const double inv_u = __longlong_as_double((long long)(1023 - m.e2) << 52);
            int S0 = 0, S1 = 0;
#pragma unroll 4
            for (uint32_t i = 0; i < 32u; i++) {
                if (i < have) {
                    const float mm = mine[4u * i + 3u], vv = mine[4u * i + qv];
                    const double tt = (double)(q == 3u ? mm : __fmul_rn(mm, vv)) * inv_u;   // exact: scaling by a power of two
                    const bool okm = fabs(tt) < 33554432.0;                      // a summand this large leaves the binade anyway
                    if (!okm) m.usable = 0;
                    const double fl = floor(tt);
                    const double fr = tt - fl;                                   // exact
                    const int fll = okm ? (int)fl : 0;
                    const int bi = fll + (fr > 0.5 ? 1 : 0);
                    const bool tie = fr == 0.5;
                    S0 += bi + ((tie && ((S0 + fll) & 1)) ? 1 : 0);              // entering even: k + S0 + floor odd -> round up
                    S1 += bi + ((tie && ((1 + S1 + fll) & 1)) ? 1 : 0);          // entering odd
                    m.mn0 = min(m.mn0, S0); m.mx0 = max(m.mx0, S0);
                    m.mn1 = min(m.mn1, S1); m.mx1 = max(m.mx1, S1);
                    if (max(abs(S0), abs(S1)) > XL_LIM) m.usable = 0;
                }
            }
            m.a0 = S0; m.a1 = S1;
            if (max(max(abs(m.mn0), abs(m.mx0)), max(abs(m.mn1), abs(m.mx1))) > XL_LIM && have) m.usable = 0;
            if (!m.usable) { m.a0 = m.a1 = 0; m.mn0 = m.mn1 = 0; m.mx0 = m.mx1 = 0; }
            // ordered reduction over the eight lanes of the group: lane k absorbs lane k + off
#pragma unroll
            for (int off = 1; off < 8; off <<= 1) {
                const XlMap o = xl_shfl_down(m, off);
                if ((sub & (2u * off - 1u)) == 0u) m = xl_combine(m, o);
            }
        }
        if (sub == 0u) maps_all[slot * 4u + q] = m;
    }
    }
}

struct XlWin { XlMap cur, nxt; };
__device__ __forceinline__ XlMap xl_load(const XlMap* __restrict__ maps, uint32_t c, uint32_t nchunks, uint32_t q) {
    if (c < nchunks) return maps[(size_t)c * 4u + q];
    return xl_identity(0, 0);
}
// the window of 32 maps moves on by nv chunks: lane i takes what lane i + nv held, the tail comes from the next window
__device__ __forceinline__ int xl_shift(int cur, int nxt, uint32_t nv, uint32_t lane) {
    const int a = __shfl_sync(0xFFFFFFFFu, cur, (int)((lane + nv) & 31u)), b = __shfl_sync(0xFFFFFFFFu, nxt, (int)((lane + nv) & 31u));
    return lane + nv < 32u ? a : b;
}

// ... and the window behind it: lane i takes what lane i + nv held, the last nv lanes take the maps they fetched themselves
__device__ __forceinline__ int xl_slide(int v, int fresh_own, uint32_t nv, uint32_t lane) {
    const int a = __shfl_sync(0xFFFFFFFFu, v, (int)((lane + nv) & 31u));
    return lane + nv < 32u ? a : fresh_own;
}

__global__ void __launch_bounds__(128) xl_chain_kernel(TreeDev t, int d0, uint32_t* __restrict__ gate) {
    __shared__ __align__(16) float buf[4][XL_CHUNK];
    __shared__ float s_mass;
    if (gate && (gate[0] == 0u || gate[1] != 0u)) return;
    const int d = d0 + (int)blockIdx.y;                 // grid.y: levels
    if (d > XL_DEPTH_MAX) return;
    const uint32_t area = (uint32_t)(d % SFE_SIDE_STREAMS);
    const XlNode* __restrict__ list = reinterpret_cast<const XlNode*>(t.xl_list) + (size_t)area * (t.xl_max + 1u);
    if (blockIdx.x >= list[t.xl_max].n) return;
    const XlNode nd = list[blockIdx.x];
    const uint32_t lane = threadIdx.x & 31u, q = threadIdx.x >> 5;   // one warp per component
    const SlotRec* __restrict__ src = t.srec[d] + nd.first;
    const XlMap* __restrict__ maps = reinterpret_cast<const XlMap*>(t.xl_map) + ((size_t)area * t.xl_slots + nd.slot0) * 4u;
    const int B0 = (1 << 23) + 1, B1 = (1 << 24) - 1;
    float s = 0.f;
    uint32_t c = 0u;
    // `cur`: the maps of chunks c .. c+31 (lane i: chunk c+i); `nxt`: those of chunks c+32 .. c+63, in flight while cur is used
    XlMap cur = xl_load(maps, c + lane, nd.nchunks, q);
    XlMap nxt = xl_load(maps, c + 32u + lane, nd.nchunks, q);
    // Members for the member-by-member fold of chunk ft_c, fetched as soon as the window shows a chunk whose map cannot
    // apply: {m, coordinate} as loaded - nothing touches the registers before the fold stages them, so the loads stay in
    // flight while the warp goes on (a product computed at fetch time would wait for them right there).  `fx`: the same for
    // a chunk whose map was expected to apply and did not (its own registers, so that it never waits for loads into ft).
    float ft_m[XL_PER_LANE], ft_v[XL_PER_LANE], fx_m[XL_PER_LANE], fx_v[XL_PER_LANE];
    uint32_t ft_c = 0xFFFFFFFFu;
    const uint32_t qv = q < 3u ? q : 3u;
    auto fetch = [&](float* fm, float* fv, uint32_t ch) {
#pragma unroll
        for (uint32_t i = 0; i < (uint32_t)XL_PER_LANE; i++) {
            const uint32_t j = ch * XL_CHUNK + i * 32u + lane;
            const float* f = reinterpret_cast<const float*>(&src[j < nd.cnt ? j : 0u].pm);
            fm[i] = j < nd.cnt ? f[3] : 0.f;                        // pads are +0: s + 0 == s
            fv[i] = f[qv];
        }
    };
    // A component whose running sum hovers around zero (a symmetric cloud: the y sums of a disc galaxy at C5) changes binade
    // all the time, and many of its maps do not apply.  After two failures close to each other the warp therefore goes
    // chunk by chunk for a while (`slow`: no composition of 32 maps that would be thrown away), and returns to windows of
    // 32 after eight chunks in a row whose maps applied.  (A sum that only grows fails at its binade crossings - about one
    // chunk per doubling - and never enters this mode.)
    uint32_t slow = 0u, last_fail = 0xC0000000u;
#ifdef SFE_XL_DEBUG
    long long dbg_t0 = clock64(), dbg_fold = 0, dbg_fetch = 0, dbg_top = 0, dbg_valid = 0, dbg_adv = 0; uint32_t dbg_fail = 0u, dbg_iter = 0u, dbg_slowit = 0u;
#endif
    XlMap pend = xl_identity(0, 0);
    uint32_t pend_lo = 32u;
    uint32_t base = 0u;                                              // chunk of lane 0 of `cur` (== c outside the slow mode)
    // slides the two windows on by nv chunks (base += nv); the maps that enter `nxt` at its tail are fetched now and merged
    // one step later (`pend`, lanes from pend_lo on), so that nothing waits for the load it has just issued
    auto advance = [&](uint32_t nv) {
        base += nv;
        if (pend_lo < 32u) {
            const bool take = lane >= pend_lo;
            nxt.a0 = take ? pend.a0 : nxt.a0; nxt.a1 = take ? pend.a1 : nxt.a1;
            nxt.mn0 = take ? pend.mn0 : nxt.mn0; nxt.mx0 = take ? pend.mx0 : nxt.mx0;
            nxt.mn1 = take ? pend.mn1 : nxt.mn1; nxt.mx1 = take ? pend.mx1 : nxt.mx1;
            nxt.e2 = take ? pend.e2 : nxt.e2; nxt.usable = take ? pend.usable : nxt.usable;
        }
        if (nv == 32u) cur = nxt;
        else {
            cur.a0 = xl_shift(cur.a0, nxt.a0, nv, lane); cur.a1 = xl_shift(cur.a1, nxt.a1, nv, lane);
            cur.mn0 = xl_shift(cur.mn0, nxt.mn0, nv, lane); cur.mx0 = xl_shift(cur.mx0, nxt.mx0, nv, lane);
            cur.mn1 = xl_shift(cur.mn1, nxt.mn1, nv, lane); cur.mx1 = xl_shift(cur.mx1, nxt.mx1, nv, lane);
            cur.e2 = xl_shift(cur.e2, nxt.e2, nv, lane); cur.usable = xl_shift(cur.usable, nxt.usable, nv, lane);
            nxt.a0 = xl_slide(nxt.a0, 0, nv, lane); nxt.a1 = xl_slide(nxt.a1, 0, nv, lane);
            nxt.mn0 = xl_slide(nxt.mn0, 0, nv, lane); nxt.mx0 = xl_slide(nxt.mx0, 0, nv, lane);
            nxt.mn1 = xl_slide(nxt.mn1, 0, nv, lane); nxt.mx1 = xl_slide(nxt.mx1, 0, nv, lane);
            nxt.e2 = xl_slide(nxt.e2, 0, nv, lane); nxt.usable = xl_slide(nxt.usable, 0, nv, lane);
        }
        pend_lo = 32u - nv;
        pend = xl_identity(0, 0);
        if (lane >= pend_lo) pend = xl_load(maps, base + 32u + lane, nd.nchunks, q);
    };
    while (c < nd.nchunks) {
        // p: position of chunk c in the window.  Outside the slow mode the window starts at c (p == 0); inside it the window
        // stays where it is and the warp steps through it (no slide, no load per chunk), a window at a time
#ifdef SFE_XL_DEBUG
        long long dbg_i0 = clock64();
#endif
        const uint32_t p = c - base;
        const int e2 = __shfl_sync(0xFFFFFFFFu, cur.e2, (int)p);
        const uint32_t okmask = __ballot_sync(0xFFFFFFFFu, cur.usable != 0 && cur.e2 == e2) | ((1u << p) - 1u);
        uint32_t nlead = (okmask == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs((int)~okmask) - 1u) - p;   // composable chunks from c on
        if (p + nlead < 32u && c + nlead < nd.nchunks && ft_c != c + nlead) {   // the first chunk ahead that cannot be composed
            ft_c = c + nlead;
            fetch(ft_m, ft_v, ft_c);
        }
#ifdef SFE_XL_DEBUG
        long long dbg_i1 = clock64(); dbg_top += dbg_i1 - dbg_i0;
#endif
        uint32_t nv = 0u;
        if (slow) {
            if (nlead > 0u) {                                        // the map of chunk c alone, held by lane p
                const int a0 = __shfl_sync(0xFFFFFFFFu, cur.a0, (int)p), a1 = __shfl_sync(0xFFFFFFFFu, cur.a1, (int)p);
                const int mn0 = __shfl_sync(0xFFFFFFFFu, cur.mn0, (int)p), mx0 = __shfl_sync(0xFFFFFFFFu, cur.mx0, (int)p);
                const int mn1 = __shfl_sync(0xFFFFFFFFu, cur.mn1, (int)p), mx1 = __shfl_sync(0xFFFFFFFFu, cur.mx1, (int)p);
                const double u = __longlong_as_double((long long)(1023 + e2) << 52);
                const double kd = (double)s * __longlong_as_double((long long)(1023 - e2) << 52);   // exact
                const bool small = fabs(kd) < 33554432.0;
                const int k = small ? (int)kd : 0;
                const int par = k & 1;
                const int lo = k + (par ? mn1 : mn0), hi = k + (par ? mx1 : mx0);
                const bool valid = small && (double)k == kd &&
                                   ((k >= B0 && k <= B1 && lo >= B0 && hi <= B1) || (k <= -B0 && k >= -B1 && hi <= -B0 && lo >= -B1));
                if (valid) { s = (float)((double)(k + (par ? a1 : a0)) * u); nv = 1u; slow--; }
            }
        } else if (nlead > 0u) {                                     // warp-uniform; p == 0 here
            // compose in order: lane i holds chunks c .. c+i (meaningful below nlead)
            XlMap m = cur;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const XlMap o = xl_shfl_up(m, off);
                if (lane >= (uint32_t)off) m = xl_combine(o, m);
            }
            const double u = __longlong_as_double((long long)(1023 + e2) << 52);
            const double kd = (double)s * __longlong_as_double((long long)(1023 - e2) << 52);   // exact
            const bool small = fabs(kd) < 33554432.0;
            const int k = small ? (int)kd : 0;
            const int par = k & 1;
            const int lo = k + (par ? m.mn1 : m.mn0), hi = k + (par ? m.mx1 : m.mx0);
            const bool valid = lane < nlead && m.usable != 0 && small && (double)k == kd &&
                               ((k >= B0 && k <= B1 && lo >= B0 && hi <= B1) || (k <= -B0 && k >= -B1 && hi <= -B0 && lo >= -B1));
            const uint32_t vmask = __ballot_sync(0xFFFFFFFFu, valid);
            nv = vmask == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs((int)~vmask) - 1u;
            if (nv > 0u) {
                const int kn = __shfl_sync(0xFFFFFFFFu, k + (par ? m.a1 : m.a0), (int)nv - 1);
                s = (float)((double)kn * u);
            }
        }
#ifdef SFE_XL_DEBUG
        long long dbg_i2 = clock64(); dbg_valid += dbg_i2 - dbg_i1;
#endif
        if (nv == 0u) {
            // member-by-member fold of chunk c (tree.c:166-169): the warp stages the summands, lane 0 adds them up
#ifdef SFE_XL_DEBUG
            long long dbg_a = clock64(); dbg_fail++;
#endif
            if (ft_c == c) {
#pragma unroll
                for (uint32_t i = 0; i < (uint32_t)XL_PER_LANE; i++) buf[q][i * 32u + lane] = q == 3u ? ft_m[i] : __fmul_rn(ft_m[i], ft_v[i]);
                ft_c = 0xFFFFFFFFu;
            } else {
                fetch(fx_m, fx_v, c);
#pragma unroll
                for (uint32_t i = 0; i < (uint32_t)XL_PER_LANE; i++) buf[q][i * 32u + lane] = q == 3u ? fx_m[i] : __fmul_rn(fx_m[i], fx_v[i]);
            }
#ifdef SFE_XL_DEBUG
            __syncwarp(); long long dbg_b = clock64(); dbg_fetch += dbg_b - dbg_a;
#endif
            __syncwarp();
            if (lane == 0u) {
                const float4* row = reinterpret_cast<const float4*>(&buf[q][0]);
#pragma unroll 8
                for (int v4 = 0; v4 < XL_CHUNK / 4; v4++) {
                    const float4 v = row[v4];
                    s = __fadd_rn(s, v.x); s = __fadd_rn(s, v.y); s = __fadd_rn(s, v.z); s = __fadd_rn(s, v.w);
                }
            }
            s = __shfl_sync(0xFFFFFFFFu, s, 0);
            __syncwarp();
#ifdef SFE_XL_DEBUG
            dbg_fold += clock64() - dbg_b;
#endif
            nv = 1u;
            if (c - last_fail < 16u) slow = 8u;                      // failures come in clusters: go chunk by chunk for a while
            last_fail = c;
        }
#ifdef SFE_XL_DEBUG
        dbg_iter++; if (slow) dbg_slowit++;
#endif
        c += nv;
#ifdef SFE_XL_DEBUG
        long long dbg_i3 = clock64();
#endif
        // the windows move when the slow mode has used one up, when it ends, and after every step outside it
        if (slow) { if (c - base == 32u) advance(32u); }
        else if (c != base) advance(c - base);
#ifdef SFE_XL_DEBUG
        dbg_adv += clock64() - dbg_i3;
#endif
    }
#ifdef SFE_XL_DEBUG
    if (lane == 0u && nd.nchunks > 2000u)
        printf("xl_chain d=%d node=%u comp=%u chunks=%u iters=%u slow_iters=%u fails=%u cycles=%lld fold=%lld fetch+stage=%lld top=%lld valid=%lld adv=%lld\n", d, nd.n, q, nd.nchunks,
               dbg_iter, dbg_slowit, dbg_fail, clock64() - dbg_t0, dbg_fold, dbg_fetch, dbg_top, dbg_valid, dbg_adv);
#endif
    if (threadIdx.x == 96u) s_mass = s;
    __syncthreads();
    if (lane == 0u) {
        const float mass = s_mass;
        const float inv = 0.f < mass ? __fdiv_rn(1.0f, mass) : 1.0f;   // tree.c:303-304
        reinterpret_cast<float*>(&t.rec1[nd.n])[q] = q < 3u ? __fmul_rn(s, inv) : s;
    }
}

}  // namespace sfe
#include "sfe_top.cuh"
namespace sfe {

// All fold kernels of the levels d0..d1 (at most SFE_SIDE_STREAMS of them: the chunked fold keeps one scratch area per level
// modulo SFE_SIDE_STREAMS).  `pre` / `pre2` (optional) are recorded in front of the serial chains of the very long nodes and
// behind the folds of the shorter ones: the top-level path lets subtree_kernel - which fills every SM for its whole
// duration - wait for the parallel part of the folds and run beside the chains (a few small CTAs that are resident by then).
static int launch_moments(Ctx* c, const TreeDev& dv, int d0, int d1, uint32_t* gate, bool single, cudaStream_t s, const char* fn,
                          cudaEvent_t pre = nullptr, cudaStream_t s2 = nullptr, cudaEvent_t pre2 = nullptr) {
    static bool attr_set = false;
    if (!attr_set) {
        SFE_CUDA(fn, cudaFuncSetAttribute(moments_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ML_SMEM));
        attr_set = true;
    }
    // the folds of the shorter nodes; on a stream of their own (s2) when the caller has one, beside the chunked fold
    cudaStream_t sl = s2 ? s2 : s;
    moments_long_kernel<<<single ? 1 : c->sm_count * 2, ML_THREADS, ML_SMEM, sl>>>(dv, d0, d1, gate);
    if (check_launch(fn)) return SFE_FAILURE;
    moments_short_kernel<<<single ? 1 : c->sm_count * 8, 256, 0, sl>>>(dv, d0, d1, gate);
    if (check_launch(fn)) return SFE_FAILURE;
    if (pre2) SFE_CUDA(fn, cudaEventRecord(pre2, sl));
    if (d0 <= XL_DEPTH_MAX) {
        const int dl = d1 < XL_DEPTH_MAX ? d1 : XL_DEPTH_MAX;
        const uint32_t g2 = single ? 8u : (uint32_t)c->sm_count * 4u;
        const uint32_t epoch0 = ((dv.build_id << 5) + (uint32_t)d0 + 1u) & 0x3FFFFFFFu;   // tells this pass's chunk flags from older ones
        xl_index_kernel<<<dl - d0 + 1, 1024, 0, s>>>(dv, d0, gate);
        if (check_launch(fn)) return SFE_FAILURE;
        xl_prep_kernel<<<g2, 256, 0, s>>>(dv, d0, dl, epoch0, gate);
        if (check_launch(fn)) return SFE_FAILURE;
        if (pre) SFE_CUDA(fn, cudaEventRecord(pre, s));
        xl_chain_kernel<<<dim3(dv.xl_max, dl - d0 + 1), 128, 0, s>>>(dv, d0, gate);
        if (check_launch(fn)) return SFE_FAILURE;
    } else if (pre) SFE_CUDA(fn, cudaEventRecord(pre, s));
    return SFE_SUCCESS;
}

// ============================================================================ finalize: walk records
__global__ void finalize_kernel(TreeDev t, uint32_t* __restrict__ gate /* null: all nodes; else root only, gated */) {
    if (gate && (gate[0] == 0u || gate[1] != 0u)) return;
    const uint32_t count = gate ? 1u : t.meta[M_COUNT];
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < count; n += gridDim.x * blockDim.x) {
        const float4 b0 = *reinterpret_cast<const float4*>(&t.rec2[n]);          // lo.x lo.y lo.z hi.x
        const float4 b1 = *(reinterpret_cast<const float4*>(&t.rec2[n]) + 1);    // hi.y hi.z parent -
        const float b[6] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y};
        const uint2 sg = t.seg[n];
        const uint32_t bodies = sg.y - sg.x;
        WalkRec0 r0;
        const float ex = __fsub_rn(b[3], b[0]), ey = __fsub_rn(b[4], b[1]), ez = __fsub_rn(b[5], b[2]);
        const uint32_t kind = bodies == 0u ? KIND_EMPTY : bodies == 1u ? KIND_SINGLE : bodies <= 8u ? KIND_LEAF : KIND_INTERNAL;
        // squared-distance filters in front of the exact inside-the-box test: generous by a few ulps of the
        // coordinates, because the centre / COM are themselves rounded
        const float slack = 4e-6f * (fabsf(b[0]) + fabsf(b[1]) + fabsf(b[2]) + fabsf(b[3]) + fabsf(b[4]) + fabsf(b[5]));
        const float diag = sqrtf(ex * ex + ey * ey + ez * ez);
        uint32_t child = 0u;
        if (kind == KIND_INTERNAL) {
            r0.a[0] = __fmul_rn(0.5f, __fadd_rn(b[0], b[3]));   // gravity.c:353-357
            r0.a[1] = __fmul_rn(0.5f, __fadd_rn(b[1], b[4]));
            r0.a[2] = __fmul_rn(0.5f, __fadd_rn(b[2], b[5]));
            const float w = __fmul_rn((float)(1.0 / 3.0), __fadd_rn(__fadd_rn(ex, ey), ez));   // gravity.c:368-372
            r0.a[3] = __fmul_rn(w, w);
            // conservative squared radius about the geometric centre outside of which a point cannot be inside the box
            r0.rin2 = (0.5f * diag + slack) * (0.5f * diag + slack) * 1.0001f + 1e-30f;
            const uint32_t base = t.child_base[n];
            for (uint32_t o = 8u; o > 0u; o--) if (t.counts[(size_t)n * 8u + o - 1u] > 0u) child = base + o - 1u;
        } else {
            const float4 cm = t.rec1[n];
            r0.a[0] = cm.x; r0.a[1] = cm.y; r0.a[2] = cm.z; r0.a[3] = cm.w;
            // a target that must be self-subtracted lies inside the box: along every axis no farther from the COM than the
            // box's farther face (true for any point, so it does not matter that the COM is a rounded quotient) - a cube about
            // the COM, whose half edge goes into rin2 (not squared).  What the filter saves are the scattered loads of the cells
            // (rec2) of the leaves around a target, on the L1 data pipe the walk is bound by.  [B200] EXACT walk at C2 with the
            // sphere "one box diagonal about the COM" 1.83 ms, with the sphere through the farthest corner 1.73 ms, with this
            // cube 1.69 ms (profiles/r02_walk_filter.txt)
            const float fx = fmaxf(fabsf(cm.x - b[0]), fabsf(b[3] - cm.x)), fy = fmaxf(fabsf(cm.y - b[1]), fabsf(b[4] - cm.y));
            const float fz = fmaxf(fabsf(cm.z - b[2]), fabsf(b[5] - cm.z));
            r0.rin2 = (fmaxf(fmaxf(fx, fy), fz) + slack) * 1.0001f + 1e-30f;
        }
        r0.next = t.next[n];
        r0.child = child;
        r0.kind = kind;
        if (kind == KIND_SINGLE) r0.rin2 = -1.0f;
        t.rec0[n] = r0;
    }
}

// ============================================================================ pre-order walk records (WalkRecP, sfe_tree.cuh)
// Pre-order id of node n = (non-empty nodes whose slot range starts before n's) + (ancestors of n that start at the same
// slot).  The nodes starting at one slot are a chain parent -> first non-empty child -> ..., of consecutive depths, so one
// atomicAdd per node into recp_hist[first slot] (count << 16 | depth) gives both the count and the chain's top depth, an
// exclusive scan over the slots gives the ids, and the successor of n (the first node, in pre-order, that starts where n's
// range ends) is simply the prefix at n's end slot.
__device__ __forceinline__ int recp_depth(const TreeDev& t, uint32_t n, bool& bad) {
    // every cell is the root halved `depth` times: read the depth off the extents along the root's longest axis
    const float* b = t.bounds;
    const float e0 = b[3] - b[0], e1 = b[4] - b[1], e2 = b[5] - b[2];
    const int ax = e0 >= e1 ? (e0 >= e2 ? 0 : 2) : (e1 >= e2 ? 1 : 2);
    const float er = ax == 0 ? e0 : ax == 1 ? e1 : e2;
    const float e = t.rec2[n].hi[ax] - t.rec2[n].lo[ax];
    if (!(er > 0.0f) || !(e > 0.0f)) { bad = true; return 0; }
    const float q = log2f(er / e);
    const int d = __float2int_rn(q);
    if (!(fabsf(q - (float)d) <= 0.25f) || d < 0 || d > 31) { bad = true; return 0; }
    return d;
}
__global__ void recp_hist_kernel(TreeDev t) {
    if (t.meta[M_STATUS] != 0u) return;
    const uint32_t count = t.meta[M_COUNT];
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < count; n += gridDim.x * blockDim.x) {
        const uint2 sg = t.seg[n];
        if (sg.y == sg.x) continue;
        bool bad = false;
        const int d = recp_depth(t, n, bad);
        if (bad) atomicOr(&t.recp_par[RP_BAD], 1u);
        atomicAdd(&t.recp_hist[sg.x], (1u << 16) | (uint32_t)d);
    }
}
struct RecpScanOp {
    TreeDev t;
    __device__ uint32_t n() const { return t.meta[M_STATUS] != 0u ? 0u : t.meta[M_NUM_INBOX] + 1u; }
    __device__ uint32_t load(uint32_t s) const { return t.recp_hist[s] >> 16; }
    __device__ void store(uint32_t s, uint32_t ex) const { t.recp_pref[s] = ex; }
    __device__ void total(uint32_t c) const { t.recp_par[RP_COUNT] = c; }
};
__device__ __forceinline__ void recp_fill(const TreeDev& t, uint32_t n, WalkRecP& r, uint32_t& kind, int d, uint32_t w0b, float kr) {
    const WalkRec0 r0 = t.rec0[n];
    if (r0.kind == KIND_INTERNAL) {
        const float4 cm = t.rec1[n];
        r.a[0] = cm.x; r.a[1] = cm.y; r.a[2] = cm.z; r.a[3] = cm.w; r.a[4] = r0.a[0]; r.a[5] = r0.a[1]; r.a[6] = r0.a[2];
        // is root width^2 / 4^depth good enough for this cell (and its filter radius covered)?
        bool sure = (w0b >> 23) > 2u * (uint32_t)d + 1u && (w0b >> 23) < 255u;
        const float w2a = __uint_as_float(w0b - ((uint32_t)d << 24));
        sure = sure && fabsf(w2a - r0.a[3]) <= RECP_TOL * r0.a[3] && r0.rin2 <= kr * w2a;
        kind = sure ? RECP_SURE : RECP_EXACT;
    } else {
        r.a[0] = r0.a[0]; r.a[1] = r0.a[1]; r.a[2] = r0.a[2]; r.a[3] = r0.a[3]; r.a[4] = r0.rin2; r.a[5] = 0.f; r.a[6] = 0.f;
        kind = r0.kind == KIND_SINGLE ? RECP_SINGLE : RECP_LEAF;
    }
}
__global__ void recp_build_kernel(TreeDev t) {
    if (t.meta[M_STATUS] != 0u) return;
    const uint32_t count = t.meta[M_COUNT], M = t.meta[M_NUM_INBOX];
    const WalkRec0 root = t.rec0[0];
    const uint32_t w0b = __float_as_uint(root.a[3]);
    const float kr = root.kind == KIND_INTERNAL && root.a[3] > 0.0f ? root.rin2 / root.a[3] * 1.01f : 1.0f;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < count; n += gridDim.x * blockDim.x) {
        const uint2 sg = t.seg[n];
        if (n == 0u) { t.recp_par[RP_W0] = w0b; t.recp_par[RP_KR] = __float_as_uint(kr); }
        if (sg.y == sg.x) continue;
        bool bad = false;
        const int d = recp_depth(t, n, bad);
        const uint32_t v = t.recp_hist[sg.x], k = v >> 16, dsum = v & 0xFFFFu;
        const int d0 = k ? (int)((dsum - k * (k - 1u) / 2u) / k) : 0;       // depths d0 .. d0 + k - 1
        if (k == 0u || d < d0 || d >= d0 + (int)k) bad = true;
        const uint32_t pre = t.recp_pref[sg.x] + (uint32_t)(d - d0);
        const uint32_t nxt = sg.y < M ? t.recp_pref[sg.y] : RECP_END;
        if (pre >= RECP_END || pre >= t.nodes_max || (nxt != RECP_END && nxt >= t.nodes_max)) bad = true;
        if (bad) { atomicOr(&t.recp_par[RP_BAD], 1u); continue; }
        WalkRecP r; uint32_t kind;
        recp_fill(t, n, r, kind, d, w0b, kr);
        r.w = kind << 30 | (uint32_t)d << 24 | nxt;
        uint32_t* dst = reinterpret_cast<uint32_t*>(&t.recp[pre]);
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(__float_as_uint(r.a[0])), "r"(__float_as_uint(r.a[1])),
                     "r"(__float_as_uint(r.a[2])), "r"(__float_as_uint(r.a[3])), "r"(__float_as_uint(r.a[4])), "r"(__float_as_uint(r.a[5])),
                     "r"(__float_as_uint(r.a[6])), "r"(r.w) : "memory");
        t.bfs_of[pre] = n;
    }
}
// the root's deferred fold (tree_root_moments) has just rewritten rec0[0] / rec1[0]: bring its pre-order record (id 0) up to date
__global__ void recp_root_kernel(TreeDev t, const uint32_t* __restrict__ gate) {
    if (gate[0] == 0u || gate[1] != 0u) return;
    if (t.meta[M_STATUS] != 0u || t.meta[M_NUM_INBOX] == 0u || t.recp_par[RP_BAD] != 0u) return;
    WalkRecP r; uint32_t kind;
    const uint32_t w = t.recp[0].w;
    recp_fill(t, 0u, r, kind, 0, t.recp_par[RP_W0], __uint_as_float(t.recp_par[RP_KR]));
    r.w = kind << 30 | (w & 0x3FFFFFFFu);   // (depth 0, same successor)
    t.recp[0] = r;
}
static int launch_recp(Ctx* c, const TreeDev& dv, uint32_t N, cudaStream_t s, const char* fn) {
    Tree& t = c->tree;
    SFE_CUDA(fn, cudaMemsetAsync(t.recp_hist.p, 0, ((size_t)(N ? N : 1) + 1) * 4, s));
    SFE_CUDA(fn, cudaMemsetAsync(t.recp_par.p, 0, RP_WORDS * 4, s));
    recp_hist_kernel<<<c->sm_count * 4, 256, 0, s>>>(dv);
    if (check_launch(fn)) return SFE_FAILURE;
    RecpScanOp op{dv};
    if (run_scan(c, op, N + 1u, s, fn)) return SFE_FAILURE;
    recp_build_kernel<<<c->sm_count * 4, 256, 0, s>>>(dv);
    if (check_launch(fn)) return SFE_FAILURE;
    return walk_publish_recp(c, s, fn);
}

// the reference-layout node table (tree.h:18-24) for sfe_tree_download(): assembled from the engine's own records
__global__ void table_kernel(TreeDev t) {
    const uint32_t count = t.meta[M_COUNT];
    const SlotRec* __restrict__ rec_final = t.srec[t.meta[M_FINAL]];
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < count; n += gridDim.x * blockDim.x) {
        const WalkRec2 cell = t.rec2[n];
        const float4 cm = t.rec1[n];
        const uint2 sg = t.seg[n];
        sfe_node_t nd;
        nd.parent = cell.parent;
        nd.bodies = sg.y - sg.x;
        const uint32_t base = t.child_base[n];
        for (uint32_t k = 0; k < 8u; k++) {
            // internal: the eight children; leaf: its members in arrival (= ascending index) order (tree.c:172-175) -
            // a leaf's slot range does not move once the leaf exists, so the final order serves every level
            if (nd.bodies > 8u) nd.branch[k] = base + k;
            else nd.branch[k] = k < nd.bodies ? rec_final[sg.x + k].idx : 0u;
        }
        for (int k = 0; k < 3; k++) { nd.bounds[k] = cell.lo[k]; nd.bounds[3 + k] = cell.hi[k]; }
        nd.params[0] = cm.x; nd.params[1] = cm.y; nd.params[2] = cm.z; nd.params[3] = cm.w;
        t.nodes[n] = nd;
    }
}

// ============================================================================ host side
static bool trace_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SFE_TRACE"); v = (e && atoi(e) > 0) ? 1 : 0; }
    return v == 1;
}

static cudaEvent_t g_trace_ev[2 * (SFE_MAX_LEVELS + 3)];
static cudaEvent_t g_trace_k[2 * (SFE_MAX_LEVELS + 3)];   // SFE_TRACE: after rank_kernel / children_kernel of every level

static uint32_t sub_max_members() {   // SFE_SUBTREE_MAX: 8 switches the subtree path off (grid-wide passes down to the leaves)
    static uint32_t v = 0;
    if (v == 0) { const char* e = getenv("SFE_SUBTREE_MAX"); v = e ? (uint32_t)atoi(e) : (uint32_t)ST_MAX; if (v < 8u) v = 8u; if (v > (uint32_t)ST_MAX) v = (uint32_t)ST_MAX; }
    return v;
}

static uint32_t top_min_particles() {   // SFE_TOP_MIN: builds of at least this many particles take the one-pass top levels (sfe_top.cuh)
    static long v = -1;
    if (v < 0) { const char* e = getenv("SFE_TOP_MIN"); v = e ? atol(e) : 65536; if (v < 0) v = 0; }
    return (uint32_t)(v > 0xFFFFFFFFl ? 0xFFFFFFFFl : v);
}

static uint32_t xl_min_members() {   // SFE_XL_MIN lowers the threshold so that small test trees exercise the chunked fold
    static uint32_t v = 0;
    if (v == 0) { const char* e = getenv("SFE_XL_MIN"); v = e ? (uint32_t)atoi(e) : 8192u; if (v < 64u) v = 64u; }
    return v;
}

constexpr uint32_t REC_SHIFT = 7u;
static TreeDev make_dev(Tree& t) {
    TreeDev d;
    memset(&d, 0, sizeof d);
    d.bounds = t.bounds_dev;
    d.nodes_max = t.nodes_max;
    d.nodes = t.nodes.as<sfe_node_t>();
    d.seg = t.seg.as<uint2>();
    d.next = t.next.as<uint32_t>();
    d.child_base = t.child_base.as<uint32_t>();
    d.counts = t.counts.as<uint32_t>();
    d.keys = t.keys.as<uint64_t>();
    d.perm_out = t.perm_out.as<uint32_t>();
    for (int k = 0; k < SFE_MAX_LEVELS + 3; k++) d.srec[k] = t.srec[k].as<SlotRec>();
    d.oct = t.oct.as<uint8_t>();
    d.rank = t.rank.as<uint32_t>();
    d.meta = t.meta.as<uint32_t>();
    d.nmid = t.nmid.as<float4>();
    // one zero-initialised block per build: [tickets | long-node counters and lists | rank_kernel descriptors |
    // children_kernel descriptors]
    const uint32_t levels = SFE_MAX_LEVELS + 2;
    const size_t n1 = t.N ? t.N : 1;
    d.xl_min = xl_min_members();
    d.xl_max = (uint32_t)(n1 / d.xl_min) + 2u;
    d.xl_slots = (uint32_t)(n1 / (uint32_t)XL_CHUNK) + 3u * d.xl_max + 8u;
    d.tdesc_stride = (uint32_t)((n1 + PT_TILE - 1) / PT_TILE + 1) * DESC_WORDS;
    d.cdesc_stride = (t.nodes_max + CH_THREADS - 1) / CH_THREADS + 1;
    uint32_t* w = t.lookback.as<uint32_t>();
    d.ticket = w;                 w += 2 * levels;
    d.sub_count = w;              w += 1;
    d.sub_ticket = w;             w += 1;
    d.xl_ticket = w;              w += levels;
    d.xl_count = w;               w += levels;
    d.ml_count = w;               w += levels;
    d.xl_nodes = w;               w += (size_t)(XL_DEPTH_MAX + 1) * d.xl_max;
    d.tdesc = w;                  w += (size_t)levels * d.tdesc_stride;
    d.cdesc = reinterpret_cast<unsigned long long*>(t.lookback.as<char>() + ((size_t)(w - t.lookback.as<uint32_t>()) * 4 + 7) / 8 * 8);
    d.xl_sum = t.xl_sum.as<double>();
    d.xl_map = t.xl_map.as<int>();
    d.xl_flag = t.xl_flag.as<uint32_t>();
    d.xl_list = t.xl_list.p;
    d.sub_max = sub_max_members();
    d.sub_list = t.sub_list.as<uint2>();
    d.top_table = t.top_table.as<uint32_t>();
    d.top_tot = t.top_cols.as<uint32_t>();
    d.top_start = d.top_tot ? d.top_tot + TOP_COLS_PAD : nullptr;
    d.top_info = d.top_tot ? d.top_tot + 2 * TOP_COLS_PAD : nullptr;
    d.top_flag = d.top_tot ? d.top_tot + 3 * TOP_COLS_PAD : nullptr;
    d.top_cand = d.top_tot ? reinterpret_cast<uint2*>(d.top_tot + 3 * TOP_COLS_PAD + 4) : nullptr;
    d.top_rank = t.top_rank.as<TopRank>();
    d.summary = t.summary_dev;
    d.build_id = t.build_id;
    // Node ids: the root is 0 and every group of eight siblings starts at an id = 1 (mod 8).  The walk records begin REC_SHIFT
    // = 7 records into their buffers, so that id 1 - and with it every sibling group - starts on a 128-byte line: the eight
    // rec0 / rec2 records of a group are two lines instead of three, their rec1 records one line instead of two
    // ([B200] EXACT walk at C2 1.708 -> 1.681 ms, FAST 1.530 -> 1.505 ms: profiles/r02_walk_filter.txt).
    d.rec0 = t.rec0.as<WalkRec0>() + REC_SHIFT;
    d.rec1 = t.rec1.as<float4>() + REC_SHIFT;
    d.rec2 = t.rec2.as<WalkRec2>() + REC_SHIFT;
    d.recp = t.recp.as<WalkRecP>();
    d.bfs_of = t.bfs_of.as<uint32_t>();
    d.recp_par = t.recp_par.as<uint32_t>();
    d.recp_hist = t.recp_hist.as<uint32_t>();
    d.recp_pref = t.recp_pref.as<uint32_t>();
    return d;
}

static size_t lookback_bytes(const TreeDev& d) {
    const size_t levels = SFE_MAX_LEVELS + 2;
    const size_t words = 2 * levels + 2 + levels + levels + levels + (size_t)(XL_DEPTH_MAX + 1) * d.xl_max + levels * d.tdesc_stride;
    return (words * 4 + 7) / 8 * 8 + levels * (size_t)d.cdesc_stride * 8;
}

TreeDev tree_dev(Ctx* c) { return make_dev(c->tree); }

// On demand (the walk asks for them when it is going to use walk_p_kernel): five launches, ~45 us at N = 2^20 - more than the
// EXACT flavour gains from them, less than the FAST flavour does (profiles/r02_walk_preorder.txt).
int tree_ensure_recp(Ctx* c, cudaStream_t s, const char* fn) {
    Tree& t = c->tree;
    if (t.recp_ready) return SFE_SUCCESS;
    if (launch_recp(c, make_dev(t), t.N, s, fn)) return SFE_FAILURE;
    t.recp_ready = true;
    return SFE_SUCCESS;
}


struct Box6 { float v[6]; };
__global__ void copy_bounds_kernel(float* __restrict__ dst, const float* __restrict__ src) { if (threadIdx.x < 6) dst[threadIdx.x] = src[threadIdx.x]; }
__global__ void set_bounds_kernel(float* __restrict__ dst, Box6 b) { if (threadIdx.x < 6) dst[threadIdx.x] = b.v[threadIdx.x]; }

int launch_tree_build(Ctx* c, const float bounds[6], const float4* pm4, uint32_t N, int depth_max,
                      uint32_t nodes_max, cudaStream_t s, const float* bounds_dev) {
    const char* fn = "sfe_tree_build";
    Tree& t = c->tree;
    t.valid = false; t.status_fetched = false; t.recp_ready = false; t.gate_hint = false;
    t.level0_range_len = 0u;
    if (depth_max < 0 || depth_max > SFE_MAX_LEVELS) { set_error(fn, "depth_max %d outside [0,%d]", depth_max, SFE_MAX_LEVELS); return SFE_FAILURE; }
    if (nodes_max < 1) { set_error(fn, "nodes_max must be at least 1"); return SFE_FAILURE; }
    t.N = N; t.nodes_max = nodes_max; t.depth_max = depth_max;
    if (t.bounds_buf.reserve(6 * sizeof(float), fn)) return SFE_FAILURE;
    if (bounds_dev) {       // the tree keeps its own copy (the root-box kernel's output is overwritten by the next box / mean radius)
        copy_bounds_kernel<<<1, 32, 0, s>>>(t.bounds_buf.as<float>(), bounds_dev);
        if (check_launch(fn)) return SFE_FAILURE;
        t.bounds_dev = t.bounds_buf.as<float>();
    } else {
        Box6 b6;
        for (int k = 0; k < 6; k++) b6.v[k] = bounds[k];
        set_bounds_kernel<<<1, 32, 0, s>>>(t.bounds_buf.as<float>(), b6);
        if (check_launch(fn)) return SFE_FAILURE;
        t.bounds_dev = t.bounds_buf.as<float>();
    }
    const size_t n1 = N ? N : 1;
    if (t.seg.reserve((size_t)nodes_max * 8, fn) ||
        t.next.reserve((size_t)nodes_max * 4, fn) || t.child_base.reserve((size_t)nodes_max * 4, fn) ||
        t.counts.reserve((size_t)nodes_max * 32, fn) || t.nmid.reserve((size_t)nodes_max * 16, fn) ||
        t.perm_out.reserve(n1 * 4, fn) ||
        t.oct.reserve(n1, fn) || t.rank.reserve(n1 * 4, fn) ||
        t.meta.reserve(M_WORDS * 4, fn) || t.root_gate.reserve(8, fn) || t.rec0.reserve(((size_t)nodes_max + REC_SHIFT) * sizeof(WalkRec0), fn) ||
        t.rec1.reserve(((size_t)nodes_max + REC_SHIFT) * 16, fn) || t.rec2.reserve(((size_t)nodes_max + REC_SHIFT) * sizeof(WalkRec2), fn) ||
        t.recp.reserve((size_t)nodes_max * sizeof(WalkRecP), fn) || t.bfs_of.reserve((size_t)nodes_max * 4, fn) ||
        t.recp_par.reserve(RP_WORDS * 4, fn) || t.recp_hist.reserve((n1 + 1) * 4, fn) || t.recp_pref.reserve((n1 + 1) * 4, fn) ||
        t.srec[0].reserve(n1 * sizeof(SlotRec), fn) || t.srec[1].reserve(n1 * sizeof(SlotRec), fn) ||
        t.sub_list.reserve((n1 / 9 + 2) * 8, fn))
        return SFE_FAILURE;
    if (!t.meta_host) SFE_CUDA(fn, cudaMallocHost(&t.meta_host, sizeof(uint32_t) * M_WORDS));
    if (!t.summary_host) {
        SFE_CUDA(fn, cudaHostAlloc(&t.summary_host, sizeof(uint32_t) * 4 * (SFE_MAX_LEVELS + 6), cudaHostAllocMapped));
        memset(t.summary_host, 0, sizeof(uint32_t) * 4 * (SFE_MAX_LEVELS + 6));
        SFE_CUDA(fn, cudaHostGetDevicePointer(&t.summary_dev, t.summary_host, 0));
    }
    if (!t.ev_init) {
        for (int k = 0; k < SFE_MAX_LEVELS + 3; k++)
            SFE_CUDA(fn, cudaEventCreateWithFlags(&t.ev_perm[k], trace_enabled() ? cudaEventDefault : cudaEventDisableTiming));
        for (auto& ev : t.ev_side) SFE_CUDA(fn, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : t.ev_pre) SFE_CUDA(fn, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : t.ev_pre2) SFE_CUDA(fn, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        if (trace_enabled()) for (auto& ev : g_trace_ev) SFE_CUDA(fn, cudaEventCreate(&ev));
        if (trace_enabled()) for (auto& ev : g_trace_k) SFE_CUDA(fn, cudaEventCreate(&ev));
        t.ev_init = true;
    }
    t.build_id++;
    if (t.build_id == 0u) t.build_id = 1u;

    // upper levels in one pass (sfe_top.cuh) when the build is large enough for it to pay
    const bool use_top = N > 0u && N >= top_min_particles() && depth_max > TOP_L && N <= (1u << 26);
    uint32_t range_len = 0u, ranges = 0u;
    if (use_top) {
        const uint32_t tiles = (N + TP_TILE - 1) / TP_TILE;
        range_len = (uint32_t)TP_TILE * ((tiles + TOP_RANGES_MAX - 1u) / TOP_RANGES_MAX);
        ranges = (N + range_len - 1u) / range_len;
        if (t.top_table.reserve((size_t)ranges * TOP_COLS * 4, fn) || t.top_cols.reserve((size_t)(3 * TOP_COLS_PAD + 4 + 2 * TOP_CAND_MAX) * 4, fn) ||
            t.top_rank.reserve(n1 * sizeof(TopRank), fn))
            return SFE_FAILURE;
        for (int k = 2; k <= TOP_L; k++) if (t.srec[k].reserve(n1 * sizeof(SlotRec), fn)) return SFE_FAILURE;
    }
    TreeDev dv = make_dev(t);
    if (t.lookback.reserve(lookback_bytes(dv), fn) || t.xl_sum.reserve((size_t)SFE_SIDE_STREAMS * dv.xl_slots * 8 * 8, fn) ||
        t.xl_map.reserve((size_t)SFE_SIDE_STREAMS * dv.xl_slots * 4 * 32, fn) ||
        t.xl_list.reserve((size_t)SFE_SIDE_STREAMS * (dv.xl_max + 1) * 32, fn))
        return SFE_FAILURE;
    if (t.xl_flag.cap < (size_t)SFE_SIDE_STREAMS * dv.xl_slots * 4) {   // fresh memory: no stale epoch may look current
        if (t.xl_flag.reserve((size_t)SFE_SIDE_STREAMS * dv.xl_slots * 4, fn)) return SFE_FAILURE;
        SFE_CUDA(fn, cudaMemsetAsync(t.xl_flag.p, 0, t.xl_flag.cap, s));
    }
    dv = make_dev(t);
    SFE_CUDA(fn, cudaMemsetAsync(t.lookback.p, 0, lookback_bytes(dv), s));
    const uint32_t pgrid = [&] { uint32_t g = (N + PT_TILE - 1) / PT_TILE; uint32_t cap = (uint32_t)c->sm_count * 6; return g < 1 ? 1u : (g < cap ? g : cap); }();
    const uint32_t cgrid = [&] { uint32_t g = (nodes_max + CH_THREADS - 1) / CH_THREADS; uint32_t cap = (uint32_t)c->sm_count * 4; return g < 1 ? 1u : (g < cap ? g : cap); }();
    const uint32_t sgrid = [&] { uint32_t g = (N + 255) / 256; uint32_t cap = (uint32_t)c->sm_count * 16; return g < 1 ? 1u : (g < cap ? g : cap); }();


    // summary[lv] = {build id, internal nodes of level lv, status, node count}, written by the device into mapped
    // host memory; the host only spins on it (no stream operation), two levels behind the device
    auto wait_summary = [&](int lv, uint32_t* internal, uint32_t* status) -> int {
        volatile uint32_t* h = t.summary_host + 4 * lv;
        uint64_t spins = 0;
        while (h[0] != t.build_id) {
            if ((++spins & 0xFFFFu) == 0u) {
                cudaError_t q = cudaStreamQuery(s);
                if (q != cudaErrorNotReady && h[0] != t.build_id) {
                    if (q == cudaSuccess) set_error(fn, "level %d never reported", lv);
                    else set_error(fn, "stream failed while building level %d: %s", lv, cudaGetErrorString(q));
                    return SFE_FAILURE;
                }
            }
        }
        *internal = h[1]; *status = h[2];
        return SFE_SUCCESS;
    };

    int first_level = 0;
    bool top_folds = false;                      // the folds of levels 1..TOP_L are on side[1] / side2[1]
    if (use_top) {
        static bool tn_attr = false;
        const size_t tn_smem = (size_t)(3 * TOP_COLS_PAD) * 4;
        if (!tn_attr) {
            SFE_CUDA(fn, cudaFuncSetAttribute(top_nodes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tn_smem));
            tn_attr = true;
        }
        SFE_CUDA(fn, cudaMemsetAsync(dv.top_flag, 0, 16, s));
        if (trace_enabled()) cudaEventRecord(t.ev_perm[0], s);
        top_rank_kernel<<<ranges, TP_THREADS, 0, s>>>(dv, pm4, N, range_len);
        if (check_launch(fn)) return SFE_FAILURE;
        top_scan_kernel<<<(TOP_COLS + 31u) / 32u, TSC_WARPS * 32, 0, s>>>(dv, ranges);
        if (check_launch(fn)) return SFE_FAILURE;
        top_nodes_kernel<<<TN_CTAS, TN_THREADS, tn_smem, s>>>(dv, depth_max, range_len);
        if (check_launch(fn)) return SFE_FAILURE;
        top_scatter_kernel<<<ranges, TP_THREADS, 0, s>>>(dv, pm4, N, range_len);
        if (check_launch(fn)) return SFE_FAILURE;
        SFE_CUDA(fn, cudaEventRecord(t.ev_perm[TOP_L], s));
        uint32_t not_applicable = 0, status = 0;
        if (wait_summary((int)TOP_SUMMARY_SLOT, &not_applicable, &status)) return SFE_FAILURE;
        if (!not_applicable) {
            {   // (written in front of the summary's build id: visible by now)
                volatile uint32_t* hh = t.summary_host + 4 * (TOP_SUMMARY_SLOT + 1u);
                const uint32_t krb = hh[1];
                float kr; memcpy(&kr, &krb, 4);
                t.gate_hint = hh[2] == 0u; t.hint_inbox = hh[0]; t.hint_kr = kr;   // (no hint with a body at the root's very centre)
            }
            first_level = TOP_L;
            t.level0_range_len = range_len;
            // the folds of levels 1 .. TOP_L: one launch of each fold kernel for all of them, the chunked fold of the very
            // long nodes on one side stream, the folds of the shorter nodes on another
            {
                cudaStream_t side = c->side[1], side2 = c->side2[1];
                SFE_CUDA(fn, cudaStreamWaitEvent(side, t.ev_perm[TOP_L], 0));
                SFE_CUDA(fn, cudaStreamWaitEvent(side2, t.ev_perm[TOP_L], 0));
                if (trace_enabled()) cudaEventRecord(g_trace_ev[2], side);
                if (launch_moments(c, dv, 1, TOP_L, nullptr, false, side, fn, t.ev_pre[1], side2, t.ev_pre2[1])) return SFE_FAILURE;
                if (trace_enabled()) cudaEventRecord(g_trace_ev[3], side);
                SFE_CUDA(fn, cudaEventRecord(t.ev_pre2[0], side));   // end of the top levels' chunked folds (their scratch areas are free again)
                top_folds = true;
            }
            if (trace_enabled()) {
                cudaStreamSynchronize(s);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, t.ev_perm[0], t.ev_perm[TOP_L]);
                fprintf(stderr, "[sfe trace] top levels 0..%d in one pass: %8.1f us\n", TOP_L - 1, ms * 1e3);
            }
        }
    }
    if (first_level == 0) {
        // level 0: in-box compaction + root
        CompactOp op{pm4, N, dv};
        if (run_scan(c, op, N, s, fn)) return SFE_FAILURE;
        root_kernel<<<1, 1, 0, s>>>(dv, depth_max);
        if (check_launch(fn)) return SFE_FAILURE;
        SFE_CUDA(fn, cudaEventRecord(t.ev_perm[0], s));
    }

    int levels_run = 0;
    int last_fold[2] = {-1, -1};                 // the last two levels whose folds went to a side stream inside the loop below
    double host_us[SFE_MAX_LEVELS + 4];
    const auto host_t0 = std::chrono::steady_clock::now();
    for (int d = first_level; d <= depth_max; d++) {
        host_us[d] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - host_t0).count();
        // level d only exists if level d-1 has internal nodes; the host learns that one level late, so the kernels
        // of at most one empty level are enqueued (they return at once)
        if (d >= 2) {
            uint32_t internal = 0, status = 0;
            if (wait_summary(d - 1, &internal, &status)) return SFE_FAILURE;
            if (internal == 0u || status != 0u) break;
        }
        // make sure perm[d+1] exists (first build at this depth only) and refresh the device view
        if (t.srec[d + 1].cap < n1 * sizeof(SlotRec)) {
            if (t.srec[d + 1].reserve(n1 * sizeof(SlotRec), fn)) return SFE_FAILURE;
            dv = make_dev(t);
        }
        // the top levels left nothing to split at level TOP_L (its summary is already there): no passes at all
        if (first_level > 0 && d == first_level) {
            uint32_t internal = 0, status = 0;
            if (wait_summary(d, &internal, &status)) return SFE_FAILURE;
            if (internal == 0u || status != 0u) break;
        }
        levels_run = d + 1;
        // moments of level d on a side stream (needs perm[d]); the root's fold is deferred
        if (d > 0 && !(first_level > 0 && d == first_level)) {
            cudaStream_t side = c->side[d % SFE_SIDE_STREAMS];
            SFE_CUDA(fn, cudaStreamWaitEvent(side, t.ev_perm[d], 0));
            if (top_folds) SFE_CUDA(fn, cudaStreamWaitEvent(side, t.ev_pre2[0], 0));   // (levels TOP_L+1.. reuse the scratch areas of the top levels)
            if (trace_enabled()) cudaEventRecord(g_trace_ev[2 * d], side);
            if (launch_moments(c, dv, d, d, nullptr, false, side, fn)) return SFE_FAILURE;
            if (trace_enabled()) cudaEventRecord(g_trace_ev[2 * d + 1], side);
            SFE_CUDA(fn, cudaEventRecord(t.ev_pre[d % SFE_SIDE_STREAMS], side));
            last_fold[1] = last_fold[0]; last_fold[0] = d;
        }
        rank_kernel<<<pgrid, PT_THREADS, 0, s>>>(dv, d);
        if (check_launch(fn)) return SFE_FAILURE;
        if (trace_enabled()) cudaEventRecord(g_trace_k[2 * d], s);
        children_kernel<<<cgrid, CH_THREADS, 0, s>>>(dv, d, depth_max);
        if (check_launch(fn)) return SFE_FAILURE;
        if (trace_enabled()) cudaEventRecord(g_trace_k[2 * d + 1], s);
        scatter_kernel<<<sgrid, 256, 0, s>>>(dv, d);
        if (check_launch(fn)) return SFE_FAILURE;
        SFE_CUDA(fn, cudaEventRecord(t.ev_perm[d + 1], s));
    }
    if (trace_enabled()) {   // SFE_TRACE=1: where does the build spend its time (device time per level vs host enqueue time)
        host_us[levels_run] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - host_t0).count();
        cudaStreamSynchronize(s);
        for (int d = first_level; d < levels_run; d++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, t.ev_perm[d], t.ev_perm[d + 1]);
            float r = 0.f, ch = 0.f, sc = 0.f;
            cudaEventElapsedTime(&r, t.ev_perm[d], g_trace_k[2 * d]);
            cudaEventElapsedTime(&ch, g_trace_k[2 * d], g_trace_k[2 * d + 1]);
            cudaEventElapsedTime(&sc, g_trace_k[2 * d + 1], t.ev_perm[d + 1]);
            fprintf(stderr, "[sfe trace] level %2d: device %8.1f us (rank %.1f children %.1f scatter %.1f)   host enqueue at %8.1f us (+%.1f)\n", d, ms * 1e3, r * 1e3, ch * 1e3, sc * 1e3, host_us[d], host_us[d + 1] - host_us[d]);
        }
    }
    // everything below the subtree roots: one CTA per root, drawn from a ticket (the final slot arrays are rewritten in
    // place range by range; the folds on the side streams read other ranges or other levels' arrays)
    if (top_folds) {
        // the parallel part of the top levels' folds goes first (subtree_kernel fills every SM for its whole duration);
        // the serial chains of the very long nodes - a few small CTAs - run beside it
        SFE_CUDA(fn, cudaStreamWaitEvent(s, t.ev_pre[1], 0));
        SFE_CUDA(fn, cudaStreamWaitEvent(s, t.ev_pre2[1], 0));
    }
    // subtree_kernel reorders the final slot array in place: the folds that read the subtree roots living there (the last
    // levels') come first
    for (int k = 0; k < 2; k++) if (last_fold[k] >= 0) SFE_CUDA(fn, cudaStreamWaitEvent(s, t.ev_pre[last_fold[k] % SFE_SIDE_STREAMS], 0));
    if (dv.sub_max > 8u) {
        static bool st_attr = false;
        if (!st_attr) {
            SFE_CUDA(fn, cudaFuncSetAttribute(subtree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StShared)));
            st_attr = true;
        }
        const uint32_t roots_max = N / 9u + 1u;
        const uint32_t sg = roots_max < (uint32_t)c->sm_count * 2u ? roots_max : (uint32_t)c->sm_count * 2u;
        if (trace_enabled()) cudaEventRecord(g_trace_ev[0], s);
        subtree_kernel<<<sg, ST_THREADS, sizeof(StShared), s>>>(dv, depth_max);
        if (check_launch(fn)) return SFE_FAILURE;
        if (trace_enabled()) {
            cudaEventRecord(g_trace_ev[1], s);
            cudaStreamSynchronize(s);
            float a = 0.f;
            cudaEventElapsedTime(&a, g_trace_ev[0], g_trace_ev[1]);
            fprintf(stderr, "[sfe trace] subtree_kernel: %8.1f us\n", a * 1e3);
        }
    }
    // join the side streams, then build the walk records
    for (int k = 0; k < SFE_SIDE_STREAMS; k++) {
        SFE_CUDA(fn, cudaEventRecord(t.ev_side[k], c->side[k]));
        SFE_CUDA(fn, cudaStreamWaitEvent(s, t.ev_side[k], 0));
    }
    if (trace_enabled()) cudaEventRecord(g_trace_ev[0], s);
    finalize_kernel<<<c->sm_count * 4, 256, 0, s>>>(dv, nullptr);
    if (check_launch(fn)) return SFE_FAILURE;
    if (trace_enabled()) {
        cudaEventRecord(g_trace_ev[1], s);
        cudaStreamSynchronize(s);
        float a = 0.f, b = 0.f;
        if (top_folds) {
            cudaEventElapsedTime(&a, t.ev_perm[0], g_trace_ev[2]);
            cudaEventElapsedTime(&b, t.ev_perm[0], g_trace_ev[3]);
            fprintf(stderr, "[sfe trace] moments levels 1..%d: start %8.1f us  end %8.1f us\n", TOP_L, a * 1e3, b * 1e3);
        }
        for (int d = first_level > 0 ? first_level + 1 : 1; d < levels_run; d++) {
            cudaEventElapsedTime(&a, t.ev_perm[0], g_trace_ev[2 * d]);
            cudaEventElapsedTime(&b, t.ev_perm[0], g_trace_ev[2 * d + 1]);
            fprintf(stderr, "[sfe trace] moments level %2d: start %8.1f us  end %8.1f us\n", d, a * 1e3, b * 1e3);
        }
        cudaEventElapsedTime(&a, t.ev_perm[0], g_trace_ev[0]);
        cudaEventElapsedTime(&b, t.ev_perm[0], g_trace_ev[1]);
        fprintf(stderr, "[sfe trace] finalize: start %8.1f us  end %8.1f us (relative to the end of the compaction)\n", a * 1e3, b * 1e3);
    }
    SFE_CUDA(fn, cudaMemcpyAsync(t.meta_host, t.meta.p, sizeof(uint32_t) * M_WORDS, cudaMemcpyDeviceToHost, s));
    t.pm4_of_build = pm4;
    SFE_CUDA(fn, cudaMemsetAsync(t.root_gate.p, 0, 8, s));
    t.valid = true;
    return SFE_SUCCESS;
}

// Deferred fold of the root's moments (N dependent adds).  Only a target that takes the ROOT as a monopole
// needs them: an out-of-box particle far enough away, any particle when theta is large, or everybody when the
// whole tree is one leaf.  `gate` = {needed, done} lives on the device; the walk sets gate[0] with a cheap
// pre-pass (root_gate_kernel) and the two kernels below return immediately when nothing needs the root.
__global__ void root_gate_kernel(TreeDev t, const float4* __restrict__ pm4, uint32_t i0, uint32_t ni, float theta,
                                 uint32_t* __restrict__ gate) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (gate[1] != 0u || gate[0] != 0u) return;
    const uint32_t M = t.meta[M_NUM_INBOX];
    if (M == 0u) return;                 // empty root: nobody evaluates it (gravity.c:324-328)
    bool need = false;
    if (k < ni) {
        if (M <= 8u) need = true;        // the root is a leaf
        else {
            const float* b = t.bounds;
            const float4 p = pm4[i0 + k];
            const float ox = __fsub_rn(__fmul_rn(0.5f, __fadd_rn(b[0], b[3])), p.x);
            const float oy = __fsub_rn(__fmul_rn(0.5f, __fadd_rn(b[1], b[4])), p.y);
            const float oz = __fsub_rn(__fmul_rn(0.5f, __fadd_rn(b[2], b[5])), p.z);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy)), __fmul_rn(oz, oz));
            const float w = __fmul_rn((float)(1.0 / 3.0),
                                      __fadd_rn(__fadd_rn(__fsub_rn(b[3], b[0]), __fsub_rn(b[4], b[1])), __fsub_rn(b[5], b[2])));
            const float ratio = 0.0f < d2 ? __fdiv_rn(__fmul_rn(w, w), d2) : 0.0f;
            need = !(theta <= ratio);
        }
    }
    if (__any_sync(0xFFFFFFFFu, need) && (threadIdx.x & 31u) == 0u) atomicOr(&gate[0], 1u);
}
__global__ void root_done_kernel(uint32_t* gate) { if (gate[0] != 0u) gate[1] = 1u; }
__global__ void root_force_kernel(uint32_t* gate) { gate[0] = 1u; }

int tree_root_moments(Ctx* c, const float4* pm4, uint32_t i0, uint32_t ni, float theta, bool force, cudaStream_t s) {
    const char* fn = "sfe_bh";
    Tree& t = c->tree;
    TreeDev dv = make_dev(t);
    uint32_t* gate = t.root_gate.as<uint32_t>();
    // Nobody accepts the ROOT as a monopole when every particle - hence every possible target - lies inside the root box, the
    // root is a cell, and a point inside the box opens it: theta <= w^2 / d^2 holds for all d^2 <= rin2 = kr w^2 when theta kr < 1.
    // The top-level pass has told the host both numbers, so the gate's nine launches (the pre-pass over the targets and eight
    // kernels that would return at once: ~28 us per walk at C2) are not made at all.
    if (!force && t.gate_hint && t.hint_inbox == t.N && t.N > 8u && theta > 0.0f && theta * t.hint_kr < 0.99f) return SFE_SUCCESS;
    if (force) root_force_kernel<<<1, 1, 0, s>>>(gate);
    else root_gate_kernel<<<(ni + 255) / 256, 256, 0, s>>>(dv, pm4, i0, ni, theta, gate);
    if (check_launch(fn)) return SFE_FAILURE;
    if (t.level0_range_len) {        // the build left the level-0 records out: write them now if the fold is going to run
        const float4* src = pm4 ? pm4 : (const float4*)t.pm4_of_build;
        top_level0_kernel<<<c->sm_count * 8, 256, 0, s>>>(dv, src, t.N, t.level0_range_len, gate);
        if (check_launch(fn)) return SFE_FAILURE;
    }
    if (launch_moments(c, dv, 0, 0, gate, true, s, fn)) return SFE_FAILURE;
    finalize_kernel<<<1, 32, 0, s>>>(dv, gate);
    if (check_launch(fn)) return SFE_FAILURE;
    if (t.recp_ready) {
        recp_root_kernel<<<1, 1, 0, s>>>(dv, gate);
        if (check_launch(fn)) return SFE_FAILURE;
    }
    root_done_kernel<<<1, 1, 0, s>>>(gate);
    return check_launch(fn);
}

int tree_status(Ctx* c, cudaStream_t s) {
    const char* fn = "tree_build";
    Tree& t = c->tree;
    if (!t.valid && !t.meta_host) { set_error(fn, "no tree has been built"); return SFE_FAILURE; }
    if (!t.status_fetched) {
        SFE_CUDA(fn, cudaStreamSynchronize(s));
        const uint32_t* mh = t.meta_host;
        t.num_nodes = mh[M_COUNT]; t.num_inbox = mh[M_NUM_INBOX]; t.depth = mh[M_DEPTH]; t.status = (int)mh[M_STATUS]; t.final_perm = (int)mh[M_FINAL];
        t.status_fetched = true;
    }
    if (t.status == 1) { t.valid = false; set_error(fn, "Exceeded maximum number of tree nodes!"); return SFE_FAILURE; }   // tree.c:181
    if (t.status == 2) { t.valid = false; set_error(fn, "Exceeded maximum tree depth!"); return SFE_FAILURE; }             // tree.c:187
    return SFE_SUCCESS;
}

}  // namespace sfe

using namespace sfe;

extern "C" int sfe_tree_download(sfe_node_t* nodes, uint32_t max_nodes, uint32_t* num_nodes) {
    const char* fn = "sfe_tree_download";
    if (ensure_init(fn)) return SFE_FAILURE;
    Ctx* c = ctx();
    Tree& t = c->tree;
    if (!t.valid) { set_error(fn, "no valid tree"); return SFE_FAILURE; }
    if (tree_root_moments(c, nullptr, 0, 0, 0.f, true, c->stream)) return SFE_FAILURE;
    if (tree_status(c, c->stream)) return SFE_FAILURE;
    if (num_nodes) *num_nodes = t.num_nodes;
    if (!nodes) return SFE_SUCCESS;
    if (max_nodes < t.num_nodes) { set_error(fn, "buffer holds %u nodes, tree has %u", max_nodes, t.num_nodes); return SFE_FAILURE; }
    if (t.nodes.reserve((size_t)t.num_nodes * sizeof(sfe_node_t), fn)) return SFE_FAILURE;
    table_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(tree_dev(c));
    if (check_launch(fn)) return SFE_FAILURE;
    if (sub_max_members() <= 8u) {     // grid-wide passes only: the engine's numbering already is the breadth-first one
        SFE_CUDA(fn, cudaMemcpyAsync(nodes, t.nodes.p, (size_t)t.num_nodes * sizeof(sfe_node_t), cudaMemcpyDeviceToHost, c->stream));
        SFE_CUDA(fn, cudaStreamSynchronize(c->stream));
        return SFE_SUCCESS;
    }
    // Node ids below the subtree roots are handed out in CTA arrival order (sfe_subtree.cuh): renumber breadth-first -
    // the children of a level allocated eight at a time in parent order, which is the canonical numbering of the
    // level-by-level construction - on the way out.  The queue holds old ids; the position in the queue is the new id.
    const uint32_t n = t.num_nodes;
    std::vector<sfe_node_t> raw(n);
    std::vector<uint32_t> queue(n);
    SFE_CUDA(fn, cudaMemcpyAsync(raw.data(), t.nodes.p, (size_t)n * sizeof(sfe_node_t), cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(fn, cudaStreamSynchronize(c->stream));
    uint32_t head = 0, tail = 0;
    queue[tail++] = 0u;
    nodes[0] = raw[0];
    nodes[0].parent = 0u;
    while (head < tail) {
        const uint32_t nid = head, oid = queue[head++];
        if (raw[oid].bodies <= 8u) continue;
        if (tail + 8u > n) { set_error(fn, "inconsistent node table (more children than nodes)"); return SFE_FAILURE; }
        for (uint32_t o = 0; o < 8u; o++) {
            const uint32_t child = raw[oid].branch[o];
            if (child >= n) { set_error(fn, "inconsistent node table (child %u of node %u outside the table)", child, oid); return SFE_FAILURE; }
            nodes[tail] = raw[child];
            nodes[tail].parent = nid;
            nodes[nid].branch[o] = tail;
            queue[tail++] = child;
        }
    }
    if (tail != n) { set_error(fn, "inconsistent node table (%u of %u nodes reachable)", tail, n); return SFE_FAILURE; }
    return SFE_SUCCESS;
}

extern "C" int sfe_tree_keys_download(uint64_t* keys, uint32_t* perm, uint32_t N, uint32_t* num_inbox) {
    const char* fn = "sfe_tree_keys_download";
    if (ensure_init(fn)) return SFE_FAILURE;
    Ctx* c = ctx();
    Tree& t = c->tree;
    if (!t.valid || t.N != N) { set_error(fn, "no valid tree for N=%u", N); return SFE_FAILURE; }
    if (tree_status(c, c->stream)) return SFE_FAILURE;
    if (num_inbox) *num_inbox = t.num_inbox;
    if (keys) {
        if (t.keys.reserve((size_t)(N ? N : 1) * 8, fn)) return SFE_FAILURE;
        if (N) {
            keys_gather_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(tree_dev(c), N);
            if (check_launch(fn)) return SFE_FAILURE;
        }
        SFE_CUDA(fn, cudaMemcpyAsync(keys, t.keys.p, (size_t)N * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    if (perm) {
        // the sort permutation = the index column of the final slot records (gathered into the build's rank scratch)
        if (t.num_inbox) {
            perm_gather_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(tree_dev(c), t.rank.as<uint32_t>());
            if (check_launch(fn)) return SFE_FAILURE;
        }
        SFE_CUDA(fn, cudaMemcpyAsync(perm, t.rank.p, (size_t)t.num_inbox * 4, cudaMemcpyDeviceToHost, c->stream));
        SFE_CUDA(fn, cudaMemcpyAsync(perm + t.num_inbox, t.perm_out.p, (size_t)(N - t.num_inbox) * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    SFE_CUDA(fn, cudaStreamSynchronize(c->stream));
    return SFE_SUCCESS;
}
