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
//             The folds run on a second stream, overlapped with the partitioning of deeper levels; the
//             root's fold (N dependent adds) is only needed by out-of-box particles and is deferred.
//
// The stable partition is slot-centric: one segmented exclusive scan over the whole slot array per level
// (8 counters per element, reset at range heads), done as tile-local scans + a scan of tile aggregates,
// so node sizes from 9 to N members go through the same three kernels.  Node numbering is breadth-first
// (children of a level allocated 8 at a time in parent order) and deterministic.
//
// Bound: HBM / L2 bandwidth - every level streams the slot arrays a few times (B_build).
#include "sfe_internal.cuh"
#include "sfe_tree.cuh"
#include "sfe_scan.cuh"
#include <chrono>
#include <stdlib.h>

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
        if (in_box(t.bounds, p)) { t.perm0[ex] = i; t.slot_node0[ex] = 0u; t.keys[i] = 0ull; t.spm4[0][ex] = p; }
        else { t.perm_out[i - ex] = i; t.keys[i] = ~0ull; }
    }
    __device__ void total(uint32_t M) const {
        // root node (tree.c:113-115) and the per-build counters
        t.meta[M_NUM_INBOX] = M;
        t.meta[M_COUNT] = 1u;
        t.meta[M_STATUS] = 0u;
        t.meta[M_DEPTH] = 0u;
        t.meta[M_LV_BEGIN + 0] = 0u;
        t.meta[M_LV_END + 0] = 1u;
        sfe_node_t root;
        root.parent = 0u; root.bodies = M;
        for (int k = 0; k < 8; k++) root.branch[k] = 0u;
        for (int k = 0; k < 6; k++) root.bounds[k] = t.bounds[k];
        for (int k = 0; k < 4; k++) root.params[k] = 0.f;
        t.nodes[0] = root;
        t.seg[0] = make_uint2(0u, M);
        t.next[0] = NODE_END;
    }
};

// ============================================================================ per level: child allocation
// element j of the scan = j-th node of level d; value = 1 if it is internal (more than 8 bodies).
struct AllocOp {
    TreeDev t; int d; int depth_max;
    __device__ uint32_t lv_begin() const { return t.meta[M_LV_BEGIN + d]; }
    __device__ uint32_t n() const { return t.meta[M_STATUS] ? 0u : t.meta[M_LV_END + d] - t.meta[M_LV_BEGIN + d]; }
    __device__ uint32_t load(uint32_t j) const { uint2 s = t.seg[lv_begin() + j]; return (s.y - s.x) > 8u ? 1u : 0u; }
    __device__ void total(uint32_t internal) const {
        const uint32_t end = t.meta[M_LV_END + d];
        uint32_t next_end = end + 8u * internal;
        if (internal > 0u) {
            // tree.c:180-189: node budget first (success iff the final count stays below max_nodes), then depth
            if (t.nodes_max <= next_end) { t.meta[M_STATUS] = 1u; next_end = end; }
            else if (depth_max <= d) { t.meta[M_STATUS] = 2u; next_end = end; }
            else t.meta[M_DEPTH] = (uint32_t)d + 1u;
        }
        t.meta[M_LV_BEGIN + d + 1] = end;
        t.meta[M_LV_END + d + 1] = next_end;
        t.meta[M_COUNT] = next_end;
    }
    __device__ void store(uint32_t j, uint32_t rank) const {
        if (t.meta[M_STATUS]) return;
        const uint32_t n = lv_begin() + j;
        const uint2 s = t.seg[n];
        const uint32_t bodies = s.y - s.x;
        sfe_node_t* nd = &t.nodes[n];
        nd->bodies = bodies;
        if (bodies > 8u) {
            const uint32_t base = t.meta[M_LV_END + d] + 8u * rank;
            t.child_base[n] = base;
            for (int o = 0; o < 8; o++) nd->branch[o] = base + o;
        } else {
            t.child_base[n] = NODE_END;
            const uint32_t* perm = t.perm[d];
            for (uint32_t b = 0; b < 8u; b++) nd->branch[b] = b < bodies ? perm[s.x + b] : 0u;   // tree.c:172-175
        }
    }
};

// ============================================================================ per level: octants + segmented scan
constexpr int PT_THREADS = 256;
constexpr int PT_ITEMS = 8;
constexpr int PT_TILE = PT_THREADS * PT_ITEMS;

struct Cnt8 {                // eight 16-bit counters (a tile holds 2048 slots)
    uint32_t w[4];
    __device__ __forceinline__ void zero() { w[0] = w[1] = w[2] = w[3] = 0u; }
    __device__ __forceinline__ void add(const Cnt8& o) { w[0] += o.w[0]; w[1] += o.w[1]; w[2] += o.w[2]; w[3] += o.w[3]; }
    __device__ __forceinline__ void inc(uint32_t o) { w[o >> 1] += 1u << ((o & 1u) * 16u); }
    __device__ __forceinline__ uint32_t get(uint32_t o) const { return (w[o >> 1] >> ((o & 1u) * 16u)) & 0xFFFFu; }
};

// Octant of a member with the reference's tie rule; also appends the key digit.
__device__ __forceinline__ uint32_t pick_octant(const sfe_node_t* nd, float4 p, bool resident) {
    const float mx = __fmul_rn(0.5f, __fadd_rn(nd->bounds[0], nd->bounds[3]));   // tree.c:159-163
    const float my = __fmul_rn(0.5f, __fadd_rn(nd->bounds[1], nd->bounds[4]));
    const float mz = __fmul_rn(0.5f, __fadd_rn(nd->bounds[2], nd->bounds[5]));
    if (resident) return (mx < p.x ? 1u : 0u) + (my < p.y ? 2u : 0u) + (mz < p.z ? 4u : 0u);     // tree.c:244-247
    return (mx <= p.x ? 1u : 0u) + (my <= p.y ? 2u : 0u) + (mz <= p.z ? 4u : 0u);                 // tree.c:282-285
}

// Tile-local segmented scan shared by the two partition kernels.  On return, for each of the thread's
// PT_ITEMS consecutive slots: oct[k] (0xFF = inactive), head[k]; `before` = counters of the open range
// at the thread's first slot counted from the last head inside the tile (or from the tile start), and
// `head_before` tells whether such a head exists inside the tile.
struct TileScan {
    Cnt8 before; bool head_before;
    Cnt8 tile_trailing; bool tile_has_head;
};

__device__ __forceinline__ void tile_segmented_scan(const uint8_t* oct, const bool* head, TileScan* out,
                                                    uint32_t* sm /* 8 warps * 5 words + 8 */) {
    // thread-local
    Cnt8 v; v.zero(); bool f = false;
#pragma unroll
    for (int k = 0; k < PT_ITEMS; k++) {
        if (head[k]) { v.zero(); f = true; }
        if (oct[k] != 0xFFu) v.inc(oct[k]);
    }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // warp inclusive segmented scan of (f, v)
    Cnt8 iv = v; bool fi = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Cnt8 pv;
#pragma unroll
        for (int q = 0; q < 4; q++) pv.w[q] = __shfl_up_sync(0xFFFFFFFFu, iv.w[q], o);
        const bool pf = __shfl_up_sync(0xFFFFFFFFu, (int)fi, o) != 0;
        if (lane >= (uint32_t)o) { if (!fi) iv.add(pv); fi = fi || pf; }
    }
    // exclusive within warp
    Cnt8 ev; bool ef;
#pragma unroll
    for (int q = 0; q < 4; q++) ev.w[q] = __shfl_up_sync(0xFFFFFFFFu, iv.w[q], 1);
    ef = __shfl_up_sync(0xFFFFFFFFu, (int)fi, 1) != 0;
    if (lane == 0) { ev.zero(); ef = false; }
    if (lane == 31) {
#pragma unroll
        for (int q = 0; q < 4; q++) sm[warp * 5 + q] = iv.w[q];
        sm[warp * 5 + 4] = fi ? 1u : 0u;
    }
    __syncthreads();
    // carry from previous warps (sequential over at most 7 warps; tiny)
    Cnt8 cw; cw.zero(); bool cf = false;
    for (uint32_t w = 0; w < warp; w++) {
        Cnt8 t;
#pragma unroll
        for (int q = 0; q < 4; q++) t.w[q] = sm[w * 5 + q];
        if (sm[w * 5 + 4]) { cw = t; cf = true; } else cw.add(t);
    }
    out->before = ev; out->head_before = ef;
    if (!ef) { out->before.add(cw); out->head_before = cf; }
    // tile aggregate (same for every thread): fold all warps
    Cnt8 tt; tt.zero(); bool tf = false;
    for (uint32_t w = 0; w < (PT_THREADS >> 5); w++) {
        Cnt8 t;
#pragma unroll
        for (int q = 0; q < 4; q++) t.w[q] = sm[w * 5 + q];
        if (sm[w * 5 + 4]) { tt = t; tf = true; } else tt.add(t);
    }
    out->tile_trailing = tt; out->tile_has_head = tf;
    __syncthreads();
}

// load one tile's worth of per-slot state
__device__ __forceinline__ void load_tile_slots(const TreeDev& t, int d, uint32_t first,
                                                uint32_t M, uint8_t* oct, bool* head, uint32_t* node) {
    const float4* __restrict__ spm4 = t.spm4[d];
    const uint32_t* __restrict__ slot_node = t.slot_node[d & 1];
#pragma unroll
    for (int k = 0; k < PT_ITEMS; k++) {
        const uint32_t s = first + k;
        oct[k] = 0xFFu; head[k] = false; node[k] = NODE_END;
        if (s < M) {
            const uint32_t n = slot_node[s];
            if (n != NODE_END) {
                const uint2 sg = t.seg[n];
                if (sg.y - sg.x > 8u) {
                    node[k] = n;
                    head[k] = (s == sg.x);
                    oct[k] = (uint8_t)pick_octant(&t.nodes[n], spm4[s], s - sg.x < 8u);
                }
            }
        }
    }
}

// S1: tile aggregates
__global__ void __launch_bounds__(PT_THREADS) part_aggregate_kernel(TreeDev t, int d) {
    __shared__ uint32_t sm[48];
    if (t.meta[M_STATUS] || t.meta[M_LV_END + d + 1] == t.meta[M_LV_BEGIN + d + 1]) return;   // nothing splits
    const uint32_t M = t.meta[M_NUM_INBOX];
    const uint32_t ntiles = (M + PT_TILE - 1) / PT_TILE;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint8_t oct[PT_ITEMS]; bool head[PT_ITEMS]; uint32_t node[PT_ITEMS];
        load_tile_slots(t, d, tile * PT_TILE + threadIdx.x * PT_ITEMS, M, oct, head, node);
        TileScan ts;
        tile_segmented_scan(oct, head, &ts, sm);
        if (threadIdx.x == 0) {
            uint32_t* a = &t.tile_agg[(size_t)tile * 9];
#pragma unroll
            for (uint32_t o = 0; o < 8; o++) a[o] = ts.tile_trailing.get(o);
            a[8] = ts.tile_has_head ? 1u : 0u;
        }
    }
}

// S2: exclusive segmented scan over the tile aggregates.  One CTA; thread t owns a contiguous run of tiles,
// scans it locally, and the per-thread totals go through a warp-shuffle + shared-memory segmented scan.
constexpr int TS_THREADS = 1024;
struct SegVal {
    uint32_t v[8]; uint32_t f;
    __device__ __forceinline__ void zero() { for (int o = 0; o < 8; o++) v[o] = 0u; f = 0u; }
    // this = earlier (+) this, where `earlier` precedes in scan order
    __device__ __forceinline__ void absorb_earlier(const SegVal& e) {
        if (!f) { for (int o = 0; o < 8; o++) v[o] += e.v[o]; }
        f |= e.f;
    }
};
__device__ __forceinline__ SegVal seg_shfl_up(const SegVal& x, int delta) {
    SegVal r;
#pragma unroll
    for (int o = 0; o < 8; o++) r.v[o] = __shfl_up_sync(0xFFFFFFFFu, x.v[o], delta);
    r.f = __shfl_up_sync(0xFFFFFFFFu, x.f, delta);
    return r;
}
__global__ void __launch_bounds__(TS_THREADS) part_tilescan_kernel(TreeDev t, int d) {
    __shared__ uint32_t sm[32][9];
    if (t.meta[M_STATUS] || t.meta[M_LV_END + d + 1] == t.meta[M_LV_BEGIN + d + 1]) return;
    const uint32_t M = t.meta[M_NUM_INBOX];
    const uint32_t ntiles = (M + PT_TILE - 1) / PT_TILE;
    const uint32_t per = (ntiles + TS_THREADS - 1) / TS_THREADS;
    const uint32_t first = threadIdx.x * per, last = min(first + per, ntiles);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // thread-local aggregate of my run of tiles
    SegVal mine; mine.zero();
    for (uint32_t tile = first; tile < last; tile++) {
        const uint32_t* a = &t.tile_agg[(size_t)tile * 9];
        SegVal x;
#pragma unroll
        for (int o = 0; o < 8; o++) x.v[o] = a[o];
        x.f = a[8];
        x.absorb_earlier(mine);
        mine = x;
    }
    // inclusive scan across the warp
    SegVal inc = mine;
#pragma unroll
    for (int dl = 1; dl < 32; dl <<= 1) {
        SegVal e = seg_shfl_up(inc, dl);
        if (lane >= (uint32_t)dl) inc.absorb_earlier(e);
    }
    if (lane == 31) {
#pragma unroll
        for (int o = 0; o < 8; o++) sm[warp][o] = inc.v[o];
        sm[warp][8] = inc.f;
    }
    __syncthreads();
    // carry from earlier warps
    SegVal carry; carry.zero();
    for (uint32_t w = 0; w < warp; w++) {
        SegVal x;
#pragma unroll
        for (int o = 0; o < 8; o++) x.v[o] = sm[w][o];
        x.f = sm[w][8];
        x.absorb_earlier(carry);
        carry = x;
    }
    // exclusive value for this thread = carry (+) inclusive of the previous lane
    SegVal ex = seg_shfl_up(inc, 1);
    if (lane == 0) ex.zero();
    ex.absorb_earlier(carry);
    // rewrite my tiles with their exclusive carries
    SegVal run = ex;
    for (uint32_t tile = first; tile < last; tile++) {
        uint32_t* a = &t.tile_agg[(size_t)tile * 9];
        SegVal x;
#pragma unroll
        for (int o = 0; o < 8; o++) x.v[o] = a[o];
        x.f = a[8];
#pragma unroll
        for (int o = 0; o < 8; o++) a[o] = run.v[o];
        x.absorb_earlier(run);
        run = x;
    }
}

// S3: stable rank of every active slot inside its (node, octant) group; per-node octant counts
__global__ void __launch_bounds__(PT_THREADS) part_rank_kernel(TreeDev t, int d) {
    __shared__ uint32_t sm[48];
    if (t.meta[M_STATUS] || t.meta[M_LV_END + d + 1] == t.meta[M_LV_BEGIN + d + 1]) return;
    const uint32_t M = t.meta[M_NUM_INBOX];
    const uint32_t ntiles = (M + PT_TILE - 1) / PT_TILE;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint8_t oct[PT_ITEMS]; bool head[PT_ITEMS]; uint32_t node[PT_ITEMS];
        const uint32_t first = tile * PT_TILE + threadIdx.x * PT_ITEMS;
        load_tile_slots(t, d, first, M, oct, head, node);
        TileScan ts;
        tile_segmented_scan(oct, head, &ts, sm);
        uint32_t run[8];
        const uint32_t* carry = &t.tile_agg[(size_t)tile * 9];
#pragma unroll
        for (uint32_t o = 0; o < 8; o++) run[o] = ts.before.get(o) + (ts.head_before ? 0u : carry[o]);
#pragma unroll
        for (int k = 0; k < PT_ITEMS; k++) {
            if (head[k]) {
#pragma unroll
                for (uint32_t o = 0; o < 8; o++) run[o] = 0u;
            }
            if (oct[k] != 0xFFu) {
                const uint32_t s = first + k;
                uint32_t r = 0;
#pragma unroll
                for (uint32_t o = 0; o < 8; o++) if (o == oct[k]) { r = run[o]; run[o] = r + 1u; }
                t.rank[s] = r;
                t.oct[s] = oct[k];
                if (s + 1u == t.seg[node[k]].y) {       // last member: the running counts are the node's totals
#pragma unroll
                    for (uint32_t o = 0; o < 8; o++) t.counts[(size_t)node[k] * 8u + o] = run[o];
                }
            } else if (first + k < M) {
                t.oct[first + k] = 0xFFu;
            }
        }
    }
}

// A3: create the eight children of every internal node of level d (tree.c:191-274)
__global__ void children_kernel(TreeDev t, int d) {
    if (t.meta[M_STATUS]) return;
    const uint32_t b = t.meta[M_LV_BEGIN + d], e = t.meta[M_LV_END + d];
    for (uint32_t n = b + blockIdx.x * blockDim.x + threadIdx.x; n < e; n += gridDim.x * blockDim.x) {
        const uint32_t base = t.child_base[n];
        if (base == NODE_END) continue;
        const sfe_node_t* nd = &t.nodes[n];
        float lo[3], hi[3], mid[3];
        for (int k = 0; k < 3; k++) {
            lo[k] = nd->bounds[k]; hi[k] = nd->bounds[3 + k];
            mid[k] = __fmul_rn(0.5f, __fadd_rn(lo[k], hi[k]));
        }
        uint32_t start = t.seg[n].x;
        const uint32_t nxt = t.next[n];
        for (uint32_t o = 0; o < 8; o++) {
            const uint32_t c = base + o, cnt = t.counts[(size_t)n * 8u + o];
            sfe_node_t ch;
            ch.parent = n; ch.bodies = cnt;
            for (int k = 0; k < 8; k++) ch.branch[k] = 0u;
            for (int k = 0; k < 3; k++) {
                ch.bounds[k] = (o >> k) & 1u ? mid[k] : lo[k];          // tree.c:219-226
                ch.bounds[3 + k] = (o >> k) & 1u ? hi[k] : mid[k];
            }
            for (int k = 0; k < 4; k++) ch.params[k] = 0.f;
            t.nodes[c] = ch;
            t.seg[c] = make_uint2(start, start + cnt);
            // depth-first successor when c is not opened: the next NON-EMPTY sibling, else the parent's successor.
            // The reference only steps back to the parent at an empty node (gravity.c:324-328), so leaving empty
            // nodes out of the walk links changes neither the visit order nor the sums.
            uint32_t succ = nxt;
            for (uint32_t o2 = 7u; o2 > o; o2--) if (t.counts[(size_t)n * 8u + o2] > 0u) succ = base + o2;
            t.next[c] = succ;
            start += cnt;
        }
    }
}

// S4: scatter into the next level's slot order
__global__ void __launch_bounds__(256) part_scatter_kernel(TreeDev t, int d) {
    const uint32_t M = t.meta[M_NUM_INBOX];
    const bool split = !t.meta[M_STATUS] && t.meta[M_LV_END + d + 1] != t.meta[M_LV_BEGIN + d + 1];
    const uint32_t* __restrict__ perm = t.perm[d];
    uint32_t* __restrict__ perm_next = t.perm[d + 1];
    const uint32_t* __restrict__ slot_node = t.slot_node[d & 1];
    uint32_t* __restrict__ slot_node_next = t.slot_node[(d + 1) & 1];
    const float4* __restrict__ spm4 = t.spm4[d];
    float4* __restrict__ spm4_next = t.spm4[d + 1];
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < M; s += gridDim.x * blockDim.x) {
        const uint32_t idx = perm[s];
        uint32_t dest = s, owner = NODE_END;
        if (split) {
            const uint32_t o = t.oct[s];
            if (o != 0xFFu) {
                owner = t.child_base[slot_node[s]] + o;
                dest = t.seg[owner].x + t.rank[s];
                t.keys[idx] |= (uint64_t)o << (45 - 3 * d);
            }
        }
        perm_next[dest] = idx;
        slot_node_next[dest] = owner;
        spm4_next[dest] = spm4[s];
    }
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
constexpr uint32_t ML_MIN = 64u;         // nodes with more members than this take the pipelined kernel
constexpr size_t ML_SMEM = (size_t)ML_R * 8 * ML_B * 16 + (size_t)2 * 32 * ML_BP * 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(ML_THREADS) moments_long_kernel(TreeDev t, int d,
                                                                  uint32_t* __restrict__ gate /* {needed, done} or null */) {
    extern __shared__ __align__(16) unsigned char ml_smem[];
    float4 (*raw)[8][ML_B] = reinterpret_cast<float4 (*)[8][ML_B]>(ml_smem);                                  // [ML_R][8][ML_B]
    float (*prod)[32][ML_BP] = reinterpret_cast<float (*)[32][ML_BP]>(ml_smem + (size_t)ML_R * 8 * ML_B * 16);  // [2][32][ML_BP]
    if (gate && (gate[0] == 0u || gate[1] != 0u)) return;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t b = t.meta[M_LV_BEGIN + d], e = t.meta[M_LV_END + d];
    const float4* __restrict__ spm4 = t.spm4[d];
    for (uint32_t base = b + blockIdx.x * 8u; base < e; base += gridDim.x * 8u) {
        // every thread looks at the 8 segments of the group (uniform loads) for the trip count
        uint32_t maxcnt = 0u;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint2 sg = base + i < e ? t.seg[base + i] : make_uint2(0u, 0u);
            const uint32_t cn = sg.y - sg.x;
            maxcnt = max(maxcnt, cn > ML_MIN ? cn : 0u);
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
            const uint32_t mycount = (msg.y - msg.x) > ML_MIN ? (msg.y - msg.x) : 0u;
            const float mass = __shfl_sync(0xFFFFFFFFu, sum, lane | 3u);
            if (mycount) {
                const float inv = 0.f < mass ? __fdiv_rn(1.0f, mass) : 1.0f;   // tree.c:303-304
                t.nodes[base + q].params[c] = c < 3u ? __fmul_rn(sum, inv) : sum;
            }
        } else {
            // ---------------------------------------------------------------- producers: one node per warp
            const uint32_t node = warp - 1u;
            const uint2 sg = base + node < e ? t.seg[base + node] : make_uint2(0u, 0u);
            const uint32_t cnt = (sg.y - sg.x) > ML_MIN ? (sg.y - sg.x) : 0u;
            const float4* __restrict__ src = spm4 + sg.x;
            auto issue = [&](uint32_t k) {       // raw batch k -> ring slot k % ML_R
                const uint32_t j0 = k * ML_B;
                if (j0 < cnt) {
                    float4* dst = &raw[k % ML_R][node][0];
#pragma unroll
                    for (int h = 0; h < ML_B / 32; h++) {
                        const uint32_t j = lane + 32u * h;
                        if (j0 + j < cnt) cp_async16(dst + j, src + j0 + j);
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

__global__ void __launch_bounds__(256) moments_short_kernel(TreeDev t, int d, uint32_t* __restrict__ gate) {
    if (gate && (gate[0] == 0u || gate[1] != 0u)) return;
    const uint32_t c = threadIdx.x & 3u;
    const uint32_t b = t.meta[M_LV_BEGIN + d], e = t.meta[M_LV_END + d];
    const float4* __restrict__ spm4 = t.spm4[d];
    const uint32_t quads = (e - b);
    for (uint32_t qd = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; qd < ((quads + 7u) & ~7u); qd += (gridDim.x * blockDim.x) >> 2) {
        const uint32_t n = b + qd;
        const bool live = n < e;
        const uint2 sg = live ? t.seg[n] : make_uint2(0u, 0u);
        uint32_t cnt = sg.y - sg.x;
        if (cnt > ML_MIN) cnt = 0u;                               // the pipelined kernel owns the long nodes
        const float* __restrict__ src = reinterpret_cast<const float*>(spm4 + sg.x);
        float sum = 0.f;
#pragma unroll 4
        for (uint32_t j = 0; j < cnt; j++) {
            const float m = __ldg(src + 4u * j + 3u);
            const float p = __fmul_rn(m, __ldg(src + 4u * j + c));  // tree.c:166-168: product rounded, then added
            sum = __fadd_rn(sum, c == 3u ? m : p);                  // tree.c:169
        }
        const float mass = __shfl_sync(0xFFFFFFFFu, sum, (threadIdx.x & 31u) | 3u);
        if (cnt) {
            const float inv = 0.f < mass ? __fdiv_rn(1.0f, mass) : 1.0f;   // tree.c:303-304
            t.nodes[n].params[c] = c < 3u ? __fmul_rn(sum, inv) : sum;
        }
    }
}

static int launch_moments(Ctx* c, const TreeDev& dv, int d, uint32_t* gate, bool single, cudaStream_t s, const char* fn) {
    static bool attr_set = false;
    if (!attr_set) {
        SFE_CUDA(fn, cudaFuncSetAttribute(moments_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ML_SMEM));
        attr_set = true;
    }
    moments_long_kernel<<<single ? 1 : c->sm_count * 2, ML_THREADS, ML_SMEM, s>>>(dv, d, gate);
    if (check_launch(fn)) return SFE_FAILURE;
    moments_short_kernel<<<single ? 1 : c->sm_count * 8, 256, 0, s>>>(dv, d, gate);
    return check_launch(fn);
}

// ============================================================================ finalize: walk records
__global__ void finalize_kernel(TreeDev t, uint32_t* __restrict__ gate /* null: all nodes; else root only, gated */) {
    if (gate && (gate[0] == 0u || gate[1] != 0u)) return;
    const uint32_t count = gate ? 1u : t.meta[M_COUNT];
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < count; n += gridDim.x * blockDim.x) {
        const sfe_node_t nd = t.nodes[n];
        const float* b = nd.bounds;
        WalkRec0 r0;
        const float ex = __fsub_rn(b[3], b[0]), ey = __fsub_rn(b[4], b[1]), ez = __fsub_rn(b[5], b[2]);
        const uint32_t kind = nd.bodies == 0u ? KIND_EMPTY : nd.bodies == 1u ? KIND_SINGLE : nd.bodies <= 8u ? KIND_LEAF : KIND_INTERNAL;
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
            for (uint32_t o = 8u; o > 0u; o--) if (t.counts[(size_t)n * 8u + o - 1u] > 0u) child = nd.branch[o - 1u];
        } else {
            r0.a[0] = nd.params[0]; r0.a[1] = nd.params[1]; r0.a[2] = nd.params[2]; r0.a[3] = nd.params[3];
            // the COM lies inside the box, so does a target that must be self-subtracted: at most one diagonal apart
            r0.rin2 = (diag + slack) * (diag + slack) * 1.0001f + 1e-30f;
        }
        r0.next = t.next[n];
        r0.child_kind = (kind << 30) | child;
        r0.pad = 0.f;
        t.rec0[n] = r0;
        t.rec1[n] = make_float4(nd.params[0], nd.params[1], nd.params[2], nd.params[3]);
        WalkRec2 r2;
        r2.lo[0] = b[0]; r2.lo[1] = b[1]; r2.lo[2] = b[2]; r2.hi[0] = b[3]; r2.hi[1] = b[4]; r2.hi[2] = b[5];
        r2.pad[0] = r2.pad[1] = 0.f;
        t.rec2[n] = r2;
    }
}

// ============================================================================ host side
static bool trace_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SFE_TRACE"); v = (e && atoi(e) > 0) ? 1 : 0; }
    return v == 1;
}

static cudaEvent_t g_trace_ev[2 * (SFE_MAX_LEVELS + 3)];

static TreeDev make_dev(Tree& t) {
    TreeDev d;
    memset(&d, 0, sizeof d);
    for (int k = 0; k < 6; k++) d.bounds[k] = t.bounds[k];
    d.nodes_max = t.nodes_max;
    d.nodes = t.nodes.as<sfe_node_t>();
    d.seg = t.seg.as<uint2>();
    d.next = t.next.as<uint32_t>();
    d.child_base = t.child_base.as<uint32_t>();
    d.counts = t.counts.as<uint32_t>();
    d.keys = t.keys.as<uint64_t>();
    d.perm_out = t.perm_out.as<uint32_t>();
    for (int k = 0; k < SFE_MAX_LEVELS + 2; k++) { d.perm[k] = t.perm[k].as<uint32_t>(); d.spm4[k] = t.spm4[k].as<float4>(); }
    d.perm0 = d.perm[0];
    d.slot_node[0] = t.slot_node[0].as<uint32_t>();
    d.slot_node[1] = t.slot_node[1].as<uint32_t>();
    d.slot_node0 = d.slot_node[0];
    d.oct = t.oct.as<uint8_t>();
    d.rank = t.rank.as<uint32_t>();
    d.tile_agg = t.tile_agg.as<uint32_t>();
    d.meta = t.meta.as<uint32_t>();
    d.rec0 = t.rec0.as<WalkRec0>();
    d.rec1 = t.rec1.as<float4>();
    d.rec2 = t.rec2.as<WalkRec2>();
    return d;
}

TreeDev tree_dev(Ctx* c) { return make_dev(c->tree); }

int launch_tree_build(Ctx* c, const float bounds[6], const float4* pm4, uint32_t N, int depth_max,
                      uint32_t nodes_max, cudaStream_t s) {
    const char* fn = "sfe_tree_build";
    Tree& t = c->tree;
    t.valid = false; t.status_fetched = false;
    if (depth_max < 0 || depth_max > SFE_MAX_LEVELS) { set_error(fn, "depth_max %d outside [0,%d]", depth_max, SFE_MAX_LEVELS); return SFE_FAILURE; }
    if (nodes_max < 1) { set_error(fn, "nodes_max must be at least 1"); return SFE_FAILURE; }
    t.N = N; t.nodes_max = nodes_max; t.depth_max = depth_max;
    for (int k = 0; k < 6; k++) t.bounds[k] = bounds[k];
    const size_t n1 = N ? N : 1;
    if (t.nodes.reserve((size_t)nodes_max * sizeof(sfe_node_t), fn) || t.seg.reserve((size_t)nodes_max * 8, fn) ||
        t.next.reserve((size_t)nodes_max * 4, fn) || t.child_base.reserve((size_t)nodes_max * 4, fn) ||
        t.counts.reserve((size_t)nodes_max * 32, fn) || t.keys.reserve(n1 * 8, fn) || t.perm_out.reserve(n1 * 4, fn) ||
        t.slot_node[0].reserve(n1 * 4, fn) || t.slot_node[1].reserve(n1 * 4, fn) || t.oct.reserve(n1, fn) ||
        t.rank.reserve(n1 * 4, fn) || t.tile_agg.reserve(((n1 + PT_TILE - 1) / PT_TILE + 1) * 36, fn) ||
        t.meta.reserve(M_WORDS * 4, fn) || t.root_gate.reserve(8, fn) || t.rec0.reserve((size_t)nodes_max * sizeof(WalkRec0), fn) ||
        t.rec1.reserve((size_t)nodes_max * 16, fn) || t.rec2.reserve((size_t)nodes_max * sizeof(WalkRec2), fn) ||
        t.perm[0].reserve(n1 * 4, fn) || t.perm[1].reserve(n1 * 4, fn) || t.spm4[0].reserve(n1 * 16, fn) ||
        t.spm4[1].reserve(n1 * 16, fn))
        return SFE_FAILURE;
    if (!t.meta_host) SFE_CUDA(fn, cudaMallocHost(&t.meta_host, sizeof(uint32_t) * M_WORDS * (SFE_MAX_LEVELS + 2)));
    if (!t.ev_init) {
        for (int k = 0; k < SFE_MAX_LEVELS + 2; k++) {
            SFE_CUDA(fn, cudaEventCreateWithFlags(&t.ev_level[k], cudaEventDisableTiming));
            SFE_CUDA(fn, cudaEventCreateWithFlags(&t.ev_perm[k], trace_enabled() ? cudaEventDefault : cudaEventDisableTiming));
        }
        for (auto& ev : t.ev_side) SFE_CUDA(fn, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        if (trace_enabled()) for (auto& ev : g_trace_ev) SFE_CUDA(fn, cudaEventCreate(&ev));
        t.ev_init = true;
    }

    TreeDev dv = make_dev(t);
    const uint32_t pgrid = [&] { uint32_t g = (N + PT_TILE - 1) / PT_TILE; uint32_t cap = (uint32_t)c->sm_count * 8; return g < 1 ? 1u : (g < cap ? g : cap); }();
    const uint32_t sgrid = [&] { uint32_t g = (N + 255) / 256; uint32_t cap = (uint32_t)c->sm_count * 16; return g < 1 ? 1u : (g < cap ? g : cap); }();

    // level 0: in-box compaction + root
    {
        CompactOp op{pm4, N, dv};
        if (run_scan(c, op, N, s, fn)) return SFE_FAILURE;
    }
    SFE_CUDA(fn, cudaEventRecord(t.ev_perm[0], s));

    int levels_run = 0;
    double host_us[SFE_MAX_LEVELS + 3];
    const auto host_t0 = std::chrono::steady_clock::now();
    for (int d = 0; d <= depth_max; d++) {
        host_us[d] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - host_t0).count();
        // make sure perm[d+1] exists (first build at this depth only) and refresh the device view
        if (t.perm[d + 1].cap < n1 * 4 || t.spm4[d + 1].cap < n1 * 16) {
            if (t.perm[d + 1].reserve(n1 * 4, fn) || t.spm4[d + 1].reserve(n1 * 16, fn)) return SFE_FAILURE;
            dv = make_dev(t);
        }
        // the host learns level sizes two levels late, which keeps the queue full while it waits
        if (d >= 2) {
            SFE_CUDA(fn, cudaEventSynchronize(t.ev_level[d - 2]));
            const uint32_t* mh = t.meta_host + (size_t)(d - 2) * M_WORDS;
            const bool more = mh[M_STATUS] == 0 && mh[M_LV_END + d - 1] != mh[M_LV_BEGIN + d - 1];
            if (!more) break;   // level d-1 is empty (or the build failed): nothing left to split
        }
        levels_run = d + 1;
        // moments of level d on a side stream (needs perm[d]); the root's fold is deferred
        if (d > 0) {
            cudaStream_t side = c->side[d % SFE_SIDE_STREAMS];
            SFE_CUDA(fn, cudaStreamWaitEvent(side, t.ev_perm[d], 0));
            if (trace_enabled()) cudaEventRecord(g_trace_ev[2 * d], side);
            if (launch_moments(c, dv, d, nullptr, false, side, fn)) return SFE_FAILURE;
            if (trace_enabled()) cudaEventRecord(g_trace_ev[2 * d + 1], side);
        }
        // child allocation for level d
        {
            AllocOp op{dv, d, depth_max};
            if (run_scan(c, op, nodes_max, s, fn)) return SFE_FAILURE;
        }
        SFE_CUDA(fn, cudaMemcpyAsync(t.meta_host + (size_t)d * M_WORDS, t.meta.p, sizeof(uint32_t) * M_WORDS,
                                     cudaMemcpyDeviceToHost, s));
        SFE_CUDA(fn, cudaEventRecord(t.ev_level[d], s));
        part_aggregate_kernel<<<pgrid, PT_THREADS, 0, s>>>(dv, d);
        if (check_launch(fn)) return SFE_FAILURE;
        part_tilescan_kernel<<<1, TS_THREADS, 0, s>>>(dv, d);
        if (check_launch(fn)) return SFE_FAILURE;
        part_rank_kernel<<<pgrid, PT_THREADS, 0, s>>>(dv, d);
        if (check_launch(fn)) return SFE_FAILURE;
        children_kernel<<<c->sm_count * 2, 128, 0, s>>>(dv, d);
        if (check_launch(fn)) return SFE_FAILURE;
        part_scatter_kernel<<<sgrid, 256, 0, s>>>(dv, d);
        if (check_launch(fn)) return SFE_FAILURE;
        SFE_CUDA(fn, cudaEventRecord(t.ev_perm[d + 1], s));
    }
    t.final_perm = levels_run;   // perm[levels_run] is the last one written by a scatter
    if (trace_enabled()) {   // SFE_TRACE=1: where does the build spend its time (device time per level vs host enqueue time)
        host_us[levels_run] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - host_t0).count();
        cudaStreamSynchronize(s);
        for (int d = 0; d < levels_run; d++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, t.ev_perm[d], t.ev_perm[d + 1]);
            fprintf(stderr, "[sfe trace] level %2d: device %8.1f us   host enqueue at %8.1f us (+%.1f)\n", d, ms * 1e3, host_us[d], host_us[d + 1] - host_us[d]);
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
        for (int d = 1; d < levels_run; d++) {
            cudaEventElapsedTime(&a, t.ev_perm[0], g_trace_ev[2 * d]);
            cudaEventElapsedTime(&b, t.ev_perm[0], g_trace_ev[2 * d + 1]);
            fprintf(stderr, "[sfe trace] moments level %2d: start %8.1f us  end %8.1f us\n", d, a * 1e3, b * 1e3);
        }
        cudaEventElapsedTime(&a, t.ev_perm[0], g_trace_ev[0]);
        cudaEventElapsedTime(&b, t.ev_perm[0], g_trace_ev[1]);
        fprintf(stderr, "[sfe trace] finalize: start %8.1f us  end %8.1f us (relative to the end of the compaction)\n", a * 1e3, b * 1e3);
    }
    SFE_CUDA(fn, cudaMemcpyAsync(t.meta_host + (size_t)(SFE_MAX_LEVELS + 1) * M_WORDS, t.meta.p, sizeof(uint32_t) * M_WORDS,
                                 cudaMemcpyDeviceToHost, s));
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
    if (force) root_force_kernel<<<1, 1, 0, s>>>(gate);
    else root_gate_kernel<<<(ni + 255) / 256, 256, 0, s>>>(dv, pm4, i0, ni, theta, gate);
    if (check_launch(fn)) return SFE_FAILURE;
    if (launch_moments(c, dv, 0, gate, true, s, fn)) return SFE_FAILURE;
    finalize_kernel<<<1, 32, 0, s>>>(dv, gate);
    if (check_launch(fn)) return SFE_FAILURE;
    root_done_kernel<<<1, 1, 0, s>>>(gate);
    return check_launch(fn);
}

int tree_status(Ctx* c, cudaStream_t s) {
    const char* fn = "tree_build";
    Tree& t = c->tree;
    if (!t.valid && !t.meta_host) { set_error(fn, "no tree has been built"); return SFE_FAILURE; }
    if (!t.status_fetched) {
        SFE_CUDA(fn, cudaStreamSynchronize(s));
        const uint32_t* mh = t.meta_host + (size_t)(SFE_MAX_LEVELS + 1) * M_WORDS;
        t.num_nodes = mh[M_COUNT]; t.num_inbox = mh[M_NUM_INBOX]; t.depth = mh[M_DEPTH]; t.status = (int)mh[M_STATUS];
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
    SFE_CUDA(fn, cudaMemcpyAsync(nodes, t.nodes.p, (size_t)t.num_nodes * sizeof(sfe_node_t), cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(fn, cudaStreamSynchronize(c->stream));
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
    if (keys) SFE_CUDA(fn, cudaMemcpyAsync(keys, t.keys.p, (size_t)N * 8, cudaMemcpyDeviceToHost, c->stream));
    if (perm) {
        SFE_CUDA(fn, cudaMemcpyAsync(perm, t.perm[t.final_perm].p, (size_t)t.num_inbox * 4, cudaMemcpyDeviceToHost, c->stream));
        SFE_CUDA(fn, cudaMemcpyAsync(perm + t.num_inbox, t.perm_out.p, (size_t)(N - t.num_inbox) * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    SFE_CUDA(fn, cudaStreamSynchronize(c->stream));
    return SFE_SUCCESS;
}
