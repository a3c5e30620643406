// sfe_top.cuh - the upper TOP_L levels of the octree build in ONE pass over the particles instead of TOP_L dependent ones.
//
// The level-by-level construction (sfe_tree.cu) costs three dependent grid-wide kernels per level, 65-80 us each at
// N = 2^20 - almost all of it latency.  But the cell a particle falls into at level d is a pure function of its position
// and the root box: mid = 0.5f*(lo+hi) (tree.c:159-163) and the child bounds (tree.c:219-226) do not depend on the other
// particles.  The one exception is the tie rule (tree.c:244-247 strict `<` for the first eight members of a node,
// tree.c:282-285 `<=` for later arrivals), which only matters for a coordinate that equals a split plane EXACTLY.  So:
//
//   top_rank_kernel     per particle (index order): in-box test (tree.c:142-146), the 3*TOP_L-bit path code, and its STABLE
//                       rank among the earlier particles of its range with the same code prefix at every level 0..TOP_L -
//                       i.e. where it stands inside each of the TOP_L+1 cells on its path, in ascending particle index.
//                       A range is a contiguous piece of the particle array worked on by one CTA in tiles of 2048: ranks by
//                       warp votes (lane = slot; 6-bit digits, a lane carries the running counts of two of the 64 bins),
//                       levels 1-2 in index order, levels 3-4 after a tile-local stable partition by the first two digits
//                       in shared memory (segmented, same scheme as rank_kernel).  Output: a 16-byte rank record per
//                       particle, one histogram row (1 + 8 + 64 + 512 + 4096 counters) per range, and a note of every
//                       in-box particle that lies exactly on a split plane of the top levels (see "Ties").
//   top_scan_kernel     exclusive prefix of every histogram column over the ranges, in place, plus the column totals.
//   top_nodes_kernel    five CTAs (each does the one scan over the 4681 column totals for itself): which cells exist (the parent holds more than sub_max bodies and is split by the grid-wide
//                       passes), node ids in the breadth-first order of children_kernel, cell records (bounds, slot range,
//                       depth-first successor, counts, split planes, fold work lists, subtree roots) for levels 0..TOP_L and
//                       the bookkeeping of open_level() - everything root_kernel / children_kernel(0..TOP_L-1) leave behind.
//   top_scatter_kernel  per particle: slot at every level it takes part in (start of the cell + ranges before + own rank),
//                       written as one 32-byte record per level into srec[1..TOP_L] - what scatter_kernel(TOP_L-1) would
//                       have left (level 0's records only on demand, top_level0_kernel).  The grid-wide passes then
//                       continue at level TOP_L.
//
// Ties: the codes are computed with the rule of an ARRIVING particle (`<=`), which is what every member of a node but its
// first eight gets.  A particle that lies exactly on a split plane (at N = 2^20 about one per build does - fp32 positions
// are that coarse) is noted as a candidate together with the level; top_nodes_kernel looks up its rank in the cell
// concerned and only if it is one of the cell's first eight members - where the resident rule `<` would have sent it to
// the other side - reports "not applicable" through mapped host memory, as it does for a tree that does not fit the node
// budget or the depth limit inside the top levels.  The host then runs the level-by-level passes from level 0 (which
// handle all of that like the reference).  (If no candidate is among the first eight of its cell, all memberships - and
// with them all ranks - are those of the reference, by induction over the levels.)  The result is the same tree bit for
// bit either way.
#pragma once

namespace sfe {

constexpr int TOP_L = 4;
constexpr uint32_t TOP_O1 = 1u, TOP_O2 = 9u, TOP_O3 = 73u, TOP_O4 = 585u, TOP_COLS = 4681u;
constexpr int TP_THREADS = 256;
constexpr int TP_TILE = 2048;
constexpr int TP_WARPS = TP_THREADS / 32;
constexpr int TP_GROUPS = TP_TILE / TP_THREADS;
constexpr uint32_t TOP_RANGES_MAX = 1024u;
constexpr uint32_t TOP_SUMMARY_SLOT = SFE_MAX_LEVELS + 3;   // summary slot of the top phase: {build id, not applicable, status, -}
constexpr uint32_t TOP_NONE = 0xFFFFu;                      // code of a particle outside the root box
constexpr uint32_t TOP_CAND_MAX = 1024u;                    // particles on a split plane that one build may note

__host__ __device__ __forceinline__ uint32_t top_off(int d) { return d == 0 ? 0u : d == 1 ? TOP_O1 : d == 2 ? TOP_O2 : d == 3 ? TOP_O3 : TOP_O4; }

struct __align__(16) TopRank { uint16_t code, r0, r1, r2, r3, r4, pad0, pad1; };   // ranks relative to the start of the range

// path code of the first TOP_L levels, most significant digit first (x | y << 1 | z << 2 per level), with the rule of an
// arriving particle (tree.c:282-285); bit d of *ties is set when a coordinate equals a split plane of level d
__device__ __forceinline__ uint32_t top_code(const float* b, float4 p, uint32_t* ties) {
    float lo[3] = {b[0], b[1], b[2]}, hi[3] = {b[3], b[4], b[5]};
    const float q[3] = {p.x, p.y, p.z};
    uint32_t code = 0u, tied = 0u;
#pragma unroll
    for (int d = 0; d < TOP_L; d++) {
        uint32_t o = 0u;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float mid = __fmul_rn(0.5f, __fadd_rn(lo[k], hi[k]));   // tree.c:159-163
            tied |= mid == q[k] ? (1u << d) : 0u;
            const bool up = mid <= q[k];
            o |= up ? (1u << k) : 0u;
            lo[k] = up ? mid : lo[k];                                     // tree.c:219-226
            hi[k] = up ? hi[k] : mid;
        }
        code = (code << 3) | o;
    }
    *ties = tied;
    return code;
}

// votes of a group of 32 slots on a 6-bit digit
struct Votes6 { uint32_t ba, b0, b1, b2, b3, b4, b5; };
__device__ __forceinline__ Votes6 vote6(bool act, uint32_t dig) {
    Votes6 v;
    v.ba = __ballot_sync(0xFFFFFFFFu, act);
    v.b0 = __ballot_sync(0xFFFFFFFFu, (dig & 1u) != 0u);  v.b1 = __ballot_sync(0xFFFFFFFFu, (dig & 2u) != 0u);
    v.b2 = __ballot_sync(0xFFFFFFFFu, (dig & 4u) != 0u);  v.b3 = __ballot_sync(0xFFFFFFFFu, (dig & 8u) != 0u);
    v.b4 = __ballot_sync(0xFFFFFFFFu, (dig & 16u) != 0u); v.b5 = __ballot_sync(0xFFFFFFFFu, (dig & 32u) != 0u);
    return v;
}
__device__ __forceinline__ uint32_t mask6(const Votes6& v, uint32_t bin) {     // active lanes whose digit is `bin`
    return v.ba & ((bin & 1u) ? v.b0 : ~v.b0) & ((bin & 2u) ? v.b1 : ~v.b1) & ((bin & 4u) ? v.b2 : ~v.b2) &
           ((bin & 8u) ? v.b3 : ~v.b3) & ((bin & 16u) ? v.b4 : ~v.b4) & ((bin & 32u) ? v.b5 : ~v.b5);
}
// the masks of bins `lane` and `lane + 32` at once: lb[k] = all ones when bit k of the lane number is set
struct LaneBits { uint32_t l0, l1, l2, l3, l4; };
__device__ __forceinline__ LaneBits lane_bits(uint32_t lane) {
    LaneBits b;
    b.l0 = 0u - (lane & 1u); b.l1 = 0u - ((lane >> 1) & 1u); b.l2 = 0u - ((lane >> 2) & 1u); b.l3 = 0u - ((lane >> 3) & 1u); b.l4 = 0u - ((lane >> 4) & 1u);
    return b;
}
__device__ __forceinline__ void mask6_pair(const Votes6& v, const LaneBits& b, uint32_t* lo, uint32_t* hi) {
    const uint32_t common = v.ba & ~(v.b0 ^ b.l0) & ~(v.b1 ^ b.l1) & ~(v.b2 ^ b.l2) & ~(v.b3 ^ b.l3) & ~(v.b4 ^ b.l4);
    *lo = common & ~v.b5; *hi = common & v.b5;
}
__device__ __forceinline__ uint32_t mask3hi(const Votes6& v, uint32_t bin) {   // active lanes whose upper three bits are `bin`
    return v.ba & ((bin & 1u) ? v.b3 : ~v.b3) & ((bin & 2u) ? v.b4 : ~v.b4) & ((bin & 4u) ? v.b5 : ~v.b5);
}
// sum over aligned groups of eight lanes; every lane of a group receives the sum
__device__ __forceinline__ uint32_t sum8(uint32_t v) {
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 1); v += __shfl_xor_sync(0xFFFFFFFFu, v, 2); v += __shfl_xor_sync(0xFFFFFFFFu, v, 4);
    return v;
}
// lane o (0..7) picks the count of upper digit o out of per-lane counts of the 64 bins (lo: bins 0..31, hi: bins 32..63)
__device__ __forceinline__ uint32_t fold8(uint32_t lo, uint32_t hi, uint32_t lane) {
    const uint32_t slo = sum8(lo), shi = sum8(hi);
    const uint32_t a = __shfl_sync(0xFFFFFFFFu, slo, (int)((lane & 3u) * 8u)), b = __shfl_sync(0xFFFFFFFFu, shi, (int)((lane & 3u) * 8u));
    return (lane & 4u) ? b : a;
}

struct TopShared {
    uint32_t row[TOP_COLS + 3];          // histogram of the range so far (levels 0..4)
    uint16_t h4[4096];                   // level-4 histogram of the current tile
    uint16_t code[TP_TILE];              // by local id (= position in the tile)
    uint16_t a2[TP_TILE];                // local ids in the order of the first two digits (stable)
    uint16_t r3[TP_TILE], r4[TP_TILE];   // tile-local ranks at levels 3 and 4 by local id
    uint32_t wtr[TP_WARPS][64];          // per warp: counts of its chunk (pass B: of the segment open at its end)
    uint32_t whh[TP_WARPS];              // pass B: the chunk holds a segment head
    uint32_t lstart[65];                 // tile-local start of every two-digit bin
};

__global__ void __launch_bounds__(TP_THREADS, 4) top_rank_kernel(TreeDev t, const float4* __restrict__ pm4, uint32_t N, uint32_t range_len) {
    __shared__ TopShared S;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u, le = lt | (1u << lane);
    const LaneBits lb = lane_bits(lane);
    const uint32_t range = blockIdx.x;
    const uint32_t i_begin = range * range_len, i_end = min(N, i_begin + range_len);
    for (uint32_t k = tid; k < TOP_COLS; k += TP_THREADS) S.row[k] = 0u;
    for (uint32_t k = tid; k < 4096u; k += TP_THREADS) S.h4[k] = 0u;
    __syncthreads();
    for (uint32_t tile0 = i_begin; tile0 < i_end; tile0 += TP_TILE) {
        // ------------------------------------------------------------------ pass A: index order, digits 1-2
        uint32_t cd[TP_GROUPS];                     // code | active << 16
        const uint32_t id0 = warp * (TP_GROUPS * 32) + lane;
#pragma unroll
        for (int g = 0; g < TP_GROUPS; g++) {
            const uint32_t id = id0 + 32u * g, i = tile0 + id;
            uint32_t c = TOP_NONE;
            if (i < i_end) {
                const float4 p = pm4[i];
                if (in_box(t.bounds, p)) {
                    uint32_t ties;
                    c = top_code(t.bounds, p, &ties);
                    if (ties & 1u) {                                // on a split plane of the root: is it the root's very centre?
                        const float* b = t.bounds;                  // (d^2 = 0 there: the one in-box point that does NOT open the
                        if (p.x == __fmul_rn(0.5f, __fadd_rn(b[0], b[3])) && p.y == __fmul_rn(0.5f, __fadd_rn(b[1], b[4])) &&   // root,
                            p.z == __fmul_rn(0.5f, __fadd_rn(b[2], b[5]))) t.top_flag[2] = 1u;                                 // gravity.c:377)
                    }
                    for (; ties; ties &= ties - 1u) {              // on a split plane: top_nodes_kernel checks its rank there
                        const uint32_t k = atomicAdd(&t.top_flag[1], 1u);
                        if (k < TOP_CAND_MAX) t.top_cand[k] = make_uint2(i, (uint32_t)__ffs(ties) - 1u);
                    }
                }
            }
            cd[g] = c;
            S.code[id] = (uint16_t)c;
        }
        uint32_t tot_lo = 0u, tot_hi = 0u;
#pragma unroll
        for (int g = 0; g < TP_GROUPS; g++) {
            const Votes6 v = vote6(cd[g] != TOP_NONE, cd[g] >> 6);
            uint32_t m_lo, m_hi;
            mask6_pair(v, lb, &m_lo, &m_hi);
            tot_lo += __popc(m_lo); tot_hi += __popc(m_hi);
        }
        S.wtr[warp][lane] = tot_lo; S.wtr[warp][lane + 32u] = tot_hi;
        __syncthreads();
        uint32_t run_lo = 0u, run_hi = 0u, all_lo = 0u, all_hi = 0u;   // bins lane / lane + 32: warps before this one, all warps
#pragma unroll
        for (uint32_t w = 0; w < (uint32_t)TP_WARPS; w++) {
            const uint32_t a = S.wtr[w][lane], b = S.wtr[w][lane + 32u];
            if (w < warp) { run_lo += a; run_hi += b; }
            all_lo += a; all_hi += b;
        }
        if (warp == 0u) {                                           // tile-local starts of the 64 two-digit bins
            const uint32_t il = warp_incl_scan(all_lo), tl = __shfl_sync(0xFFFFFFFFu, il, 31);
            const uint32_t ih = warp_incl_scan(all_hi);
            S.lstart[lane] = il - all_lo; S.lstart[lane + 32u] = tl + ih - all_hi;
            if (lane == 31u) S.lstart[64] = tl + ih;
        }
        uint32_t run8 = fold8(run_lo, run_hi, lane);                // lanes 0..7: upper digit o, warps before this one
        uint32_t run0 = __shfl_sync(0xFFFFFFFFu, sum8(run8), 0);     // in-box slots of the warps before this one
        __syncthreads();
        uint32_t r01[TP_GROUPS], r2l[TP_GROUPS];                    // tile-local ranks: r0 | r1 << 16, r2
#pragma unroll
        for (int g = 0; g < TP_GROUPS; g++) {
            const bool act = cd[g] != TOP_NONE;
            const uint32_t q2 = (cd[g] >> 6) & 63u, q1 = q2 >> 3;
            const Votes6 v = vote6(act, q2);
            const uint32_t b_lo = __shfl_sync(0xFFFFFFFFu, run_lo, (int)(q2 & 31u)), b_hi = __shfl_sync(0xFFFFFFFFu, run_hi, (int)(q2 & 31u));
            const uint32_t r2 = __popc(mask6(v, q2) & lt) + (q2 < 32u ? b_lo : b_hi);
            const uint32_t r1 = __popc(mask3hi(v, q1) & lt) + __shfl_sync(0xFFFFFFFFu, run8, (int)q1);
            const uint32_t r0 = __popc(v.ba & lt) + run0;
            r01[g] = r0 | (r1 << 16); r2l[g] = r2;
            if (act) S.a2[S.lstart[q2] + r2] = (uint16_t)(id0 + 32u * g);
            uint32_t m_lo, m_hi;
            mask6_pair(v, lb, &m_lo, &m_hi);
            run_lo += __popc(m_lo); run_hi += __popc(m_hi);
            run8 += __popc(mask3hi(v, lane & 7u));
            run0 += __popc(v.ba);
        }
        __syncthreads();
        // ------------------------------------------------------------------ pass B: order of the first two digits, digits 3-4,
        //                                                                     segmented by the two-digit bin
        const uint32_t nin = S.lstart[64];
        uint32_t e[TP_GROUPS];                      // low six bits | two-digit bin << 6 | local id << 12 | head << 23 | last << 24 | active << 25
        uint32_t hm[TP_GROUPS];
        int gl = -1;
#pragma unroll
        for (int g = 0; g < TP_GROUPS; g++) {
            const uint32_t s = id0 + 32u * g;
            uint32_t x = 0u;
            if (s < nin) {
                const uint32_t id = S.a2[s], c = S.code[id], q2 = c >> 6;
                x = c | (id << 12) | (s == S.lstart[q2] ? 1u << 23 : 0u) | (s + 1u == S.lstart[q2 + 1u] ? 1u << 24 : 0u) | (1u << 25);
            }
            e[g] = x;
            hm[g] = __ballot_sync(0xFFFFFFFFu, (x >> 23) & 1u);
            if (hm[g]) gl = g;
        }
        tot_lo = 0u; tot_hi = 0u;                  // counts of the segment that is open at the end of the warp's chunk
#pragma unroll
        for (int g = 0; g < TP_GROUPS; g++) {
            if (g >= gl) {                          // uniform across the warp
                const Votes6 v = vote6((e[g] >> 25) & 1u, e[g] & 63u);
                const uint32_t from = g == gl ? (0xFFFFFFFFu << (31 - __clz(hm[g]))) : 0xFFFFFFFFu;
                uint32_t m_lo, m_hi;
                mask6_pair(v, lb, &m_lo, &m_hi);
                tot_lo += __popc(m_lo & from); tot_hi += __popc(m_hi & from);
            }
        }
        S.wtr[warp][lane] = tot_lo; S.wtr[warp][lane + 32u] = tot_hi;
        if (lane == 0u) S.whh[warp] = gl >= 0 ? 1u : 0u;
        __syncthreads();
        run_lo = 0u; run_hi = 0u;
        for (uint32_t w = 0; w < warp; w++) {
            const uint32_t a = S.wtr[w][lane], b = S.wtr[w][lane + 32u];
            if (S.whh[w] != 0u) { run_lo = a; run_hi = b; } else { run_lo += a; run_hi += b; }
        }
#pragma unroll
        for (int g = 0; g < TP_GROUPS; g++) {
            const bool act = ((e[g] >> 25) & 1u) != 0u;
            const uint32_t d34 = e[g] & 63u, d3 = d34 >> 3, id = (e[g] >> 12) & 2047u;
            const Votes6 v = vote6(act, d34);
            const uint32_t headmask = hm[g];
            const uint32_t hb = headmask & le;                       // heads at or before this lane
            const bool in_open = hb == 0u;                           // my segment began before this group
            const uint32_t from = in_open ? 0xFFFFFFFFu : (0xFFFFFFFFu << (31 - __clz(hb)));
            run8 = fold8(run_lo, run_hi, lane);
            const uint32_t b_lo = __shfl_sync(0xFFFFFFFFu, run_lo, (int)(d34 & 31u)), b_hi = __shfl_sync(0xFFFFFFFFu, run_hi, (int)(d34 & 31u));
            const uint32_t b8 = __shfl_sync(0xFFFFFFFFu, run8, (int)d3);
            const uint32_t r4 = __popc(mask6(v, d34) & lt & from) + (in_open ? (d34 < 32u ? b_lo : b_hi) : 0u);
            const uint32_t r3 = __popc(mask3hi(v, d3) & lt & from) + (in_open ? b8 : 0u);
            if (act) { S.r3[id] = (uint16_t)r3; S.r4[id] = (uint16_t)r4; }
            uint32_t m_lo, m_hi;
            mask6_pair(v, lb, &m_lo, &m_hi);
            // last member of a segment: the 64 totals of its two-digit bin
            for (uint32_t m = __ballot_sync(0xFFFFFFFFu, (e[g] >> 24) & 1u); m; m &= m - 1u) {
                const uint32_t L = (uint32_t)__ffs(m) - 1u, upto = 0xFFFFFFFFu >> (31u - L);
                const uint32_t hbl = headmask & upto;
                const uint32_t rng = hbl ? (upto & (0xFFFFFFFFu << (31 - __clz(hbl)))) : upto;
                const uint32_t q2 = (__shfl_sync(0xFFFFFFFFu, e[g], (int)L) >> 6) & 63u;
                S.h4[q2 * 64u + lane] = (uint16_t)(__popc(m_lo & rng) + (hbl ? 0u : run_lo));
                S.h4[q2 * 64u + lane + 32u] = (uint16_t)(__popc(m_hi & rng) + (hbl ? 0u : run_hi));
            }
            if (headmask) {
                const uint32_t fr = 0xFFFFFFFFu << (31 - __clz(headmask));
                run_lo = __popc(m_lo & fr); run_hi = __popc(m_hi & fr);
            } else { run_lo += __popc(m_lo); run_hi += __popc(m_hi); }
        }
        __syncthreads();
        // ------------------------------------------------------------------ rank records (relative to the range start)
#pragma unroll
        for (int g = 0; g < TP_GROUPS; g++) {
            const uint32_t id = id0 + 32u * g, i = tile0 + id;
            if (i < i_end) {
                const uint32_t c = cd[g];
                TopRank rk;
                rk.code = (uint16_t)c; rk.pad0 = 0; rk.pad1 = 0;
                rk.r0 = (uint16_t)(S.row[0] + (r01[g] & 0xFFFFu));
                if (c != TOP_NONE) {
                    rk.r1 = (uint16_t)(S.row[TOP_O1 + (c >> 9)] + (r01[g] >> 16));
                    rk.r2 = (uint16_t)(S.row[TOP_O2 + (c >> 6)] + r2l[g]);
                    rk.r3 = (uint16_t)(S.row[TOP_O3 + (c >> 3)] + S.r3[id]);
                    rk.r4 = (uint16_t)(S.row[TOP_O4 + c] + S.r4[id]);
                } else { rk.r1 = rk.r2 = rk.r3 = rk.r4 = 0; }
                *reinterpret_cast<uint4*>(&t.top_rank[i]) = *reinterpret_cast<const uint4*>(&rk);
            }
        }
        __syncthreads();
        // ------------------------------------------------------------------ the tile's histogram joins the range's
        for (uint32_t k = tid; k < 4096u; k += TP_THREADS) S.row[TOP_O4 + k] += S.h4[k];
        for (uint32_t k = tid; k < 512u; k += TP_THREADS) {
            uint32_t s = 0u;
#pragma unroll
            for (uint32_t o = 0; o < 8u; o++) s += S.h4[k * 8u + o];
            S.row[TOP_O3 + k] += s;
        }
        if (tid < 64u) S.row[TOP_O2 + tid] += S.lstart[tid + 1u] - S.lstart[tid];
        if (tid < 8u) S.row[TOP_O1 + tid] += S.lstart[8u * tid + 8u] - S.lstart[8u * tid];
        if (tid == 0u) S.row[0] += nin;
        __syncthreads();
        for (uint32_t k = tid; k < 4096u; k += TP_THREADS) S.h4[k] = 0u;
        __syncthreads();
    }
    uint32_t* __restrict__ out = t.top_table + (size_t)range * TOP_COLS;
    for (uint32_t k = tid; k < TOP_COLS; k += TP_THREADS) out[k] = S.row[k];
}

// exclusive prefix of every column over the ranges (in place) + column totals: a CTA owns 32 columns, its 32 warps
// split the ranges between them (all loads of a warp's share are in flight at once)
constexpr int TSC_WARPS = 32;
__global__ void __launch_bounds__(TSC_WARPS * 32) top_scan_kernel(TreeDev t, uint32_t R) {
    __shared__ uint32_t part[TSC_WARPS][33];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t col = blockIdx.x * 32u + lane;
    const uint32_t per = (R + TSC_WARPS - 1u) / TSC_WARPS, r0 = min(R, warp * per), r1 = min(R, r0 + per);
    uint32_t* __restrict__ tab = t.top_table;
    const bool live = col < TOP_COLS;
    uint32_t s = 0u;
    for (uint32_t rb = r0; rb < r1; rb += 16u) {
        uint32_t v[16];
#pragma unroll
        for (uint32_t j = 0; j < 16u; j++) v[j] = (live && rb + j < r1) ? tab[(size_t)(rb + j) * TOP_COLS + col] : 0u;
#pragma unroll
        for (uint32_t j = 0; j < 16u; j++) s += v[j];
    }
    part[warp][lane] = s;
    __syncthreads();
    uint32_t run = 0u, total = 0u;
#pragma unroll
    for (uint32_t w = 0; w < (uint32_t)TSC_WARPS; w++) { const uint32_t v = part[w][lane]; if (w < warp) run += v; total += v; }
    for (uint32_t rb = r0; rb < r1; rb += 16u) {
        uint32_t v[16];
#pragma unroll
        for (uint32_t j = 0; j < 16u; j++) v[j] = (live && rb + j < r1) ? tab[(size_t)(rb + j) * TOP_COLS + col] : 0u;
#pragma unroll
        for (uint32_t j = 0; j < 16u; j++) {
            if (live && rb + j < r1) tab[(size_t)(rb + j) * TOP_COLS + col] = run;
            run += v[j];
        }
    }
    if (live && warp == 0u) t.top_tot[col] = total;
}

__device__ __forceinline__ unsigned long long shfl_up_ull(unsigned long long v, int off) {
    const uint32_t lo = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)v, off), hi = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), off);
    return ((unsigned long long)hi << 32) | lo;
}

// bounds of the cell with path `q` (d digits) below the root box, and its split planes
__device__ __forceinline__ void top_cell(const float* b, uint32_t q, int d, float lo[3], float hi[3]) {
    for (int k = 0; k < 3; k++) { lo[k] = b[k]; hi[k] = b[3 + k]; }
    for (int lev = 0; lev < d; lev++) {
        const uint32_t o = (q >> (3 * (d - 1 - lev))) & 7u;
        for (int k = 0; k < 3; k++) {
            const float mid = __fmul_rn(0.5f, __fadd_rn(lo[k], hi[k]));   // tree.c:159-163
            if ((o >> k) & 1u) lo[k] = mid; else hi[k] = mid;             // tree.c:219-226
        }
    }
}

constexpr int TN_THREADS = 1024;
constexpr int TN_ITEMS = 5;                    // columns per thread (TN_THREADS * TN_ITEMS >= TOP_COLS)
constexpr int TN_CTAS = 5;                     // CTAs that share the per-cell work
constexpr uint32_t TOP_COLS_PAD = 4684u;
static_assert(TN_THREADS * TN_ITEMS >= (int)TOP_COLS, "one scan covers all columns");

__device__ __forceinline__ int top_level_of(uint32_t c) { return c >= TOP_O4 ? 4 : c >= TOP_O3 ? 3 : c >= TOP_O2 ? 2 : c >= TOP_O1 ? 1 : 0; }

// Everything about the nodes of the top levels follows from the column totals alone: a cell is split by the grid-wide
// passes iff it holds more than sub_max bodies (then so do all its ancestors), a cell has a node iff its parent is split,
// ids are handed out level by level, eight per split parent in path order (the breadth-first numbering of
// children_kernel), and a cell's first slot is the number of in-box particles with a smaller path.  So ONE scan over the
// 4681 columns (bodies in the low word, split cells in the high word) gives every start and every id, and every cell
// writes its own records without waiting for its parent's.
__global__ void __launch_bounds__(TN_THREADS) top_nodes_kernel(TreeDev t, int depth_max, uint32_t range_len) {
    extern __shared__ __align__(16) uint32_t tn_smem[];
    uint32_t* tot = tn_smem;                   // bodies of every cell of the top levels
    uint32_t* exb = tot + TOP_COLS_PAD;        // bodies in the columns before this one (all levels concatenated)
    uint32_t* exs = exb + TOP_COLS_PAD;        // split cells (low half) and subtree roots (high half) in the columns before this one
    __shared__ unsigned long long sw[33];
    __shared__ uint32_t s_lvb[TOP_L + 3], s_int[TOP_L + 2], s_bad;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    volatile uint32_t* hs = t.summary + 4 * TOP_SUMMARY_SLOT;
    for (uint32_t k = tid; k < TOP_COLS_PAD; k += TN_THREADS) tot[k] = k < TOP_COLS ? t.top_tot[k] : 0u;
    if (tid == 0) s_bad = 0u;
    __syncthreads();
    const uint32_t M = tot[0], sub_max = t.sub_max;
    // ---- one scan: bodies (low word) and split cells (high word) before every column
    {
        unsigned long long v[TN_ITEMS], sum = 0ull;
#pragma unroll
        for (int k = 0; k < TN_ITEMS; k++) {
            const uint32_t c = tid * TN_ITEMS + k;
            // a subtree root further down is only real if its parent is split, which the nesting of the counts guarantees
            // for every cell but the root's own children - and those are real as soon as the root is split
            const uint32_t cn = c < TOP_COLS ? tot[c] : 0u;
            const bool real = c == 0u || (c < TOP_COLS && tot[top_off(top_level_of(c) - 1) + ((c - top_off(top_level_of(c))) >> 3)] > sub_max);
            v[k] = c < TOP_COLS ? ((unsigned long long)((cn > sub_max ? 1u : 0u) | ((real && cn > 8u && cn <= sub_max) ? 0x10000u : 0u)) << 32) | cn : 0ull;
            sum += v[k];
        }
        unsigned long long inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = shfl_up_ull(inc, o);
            if (lane >= (uint32_t)o) inc += y;
        }
        if (lane == 31u) sw[warp] = inc;
        __syncthreads();
        if (warp == 0u) {
            const unsigned long long w = sw[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = shfl_up_ull(wi, o);
                if (lane >= (uint32_t)o) wi += y;
            }
            sw[lane] = wi - w;
            if (lane == 31u) sw[32] = wi;
        }
        __syncthreads();
        unsigned long long ex = inc - sum + sw[warp];
#pragma unroll
        for (int k = 0; k < TN_ITEMS; k++) {
            const uint32_t c = tid * TN_ITEMS + k;
            if (c < TOP_COLS_PAD) { exb[c] = (uint32_t)ex; exs[c] = (uint32_t)(ex >> 32); }
            ex += v[k];
        }
    }
    __syncthreads();
    // ---- does the one-pass construction apply?  Node budget and depth limit must not be reached inside the top levels (the
    //      level-by-level passes report those like the reference, tree.c:180-189), and no particle on a split plane may be
    //      one of the first eight members of the cell that is split there (tree.c:244-247)
    {
        const uint32_t ncand = t.top_flag[1];
        uint32_t bad = ncand > TOP_CAND_MAX ? 1u : 0u;
        for (uint32_t k = tid; k < min(ncand, TOP_CAND_MAX); k += TN_THREADS) {
            const uint2 cd = t.top_cand[k];                      // {particle, level of the cell whose split plane it is on}
            const uint4 raw = *reinterpret_cast<const uint4*>(&t.top_rank[cd.x]);
            const uint32_t code = raw.x & 0xFFFFu;
            const uint32_t rr = cd.y == 0u ? raw.x >> 16 : cd.y == 1u ? raw.y & 0xFFFFu : cd.y == 2u ? raw.y >> 16 : raw.z & 0xFFFFu;
            const uint32_t col = top_off((int)cd.y) + (code >> (3 * (TOP_L - (int)cd.y)));
            const uint32_t rank = t.top_table[(size_t)(cd.x / range_len) * TOP_COLS + col] + rr;
            if (rank < 8u && tot[col] > sub_max) bad = 1u;
        }
        const uint32_t nsplit = exs[TOP_COLS] & 0xFFFFu;         // split cells of all top levels
        if ((unsigned long long)1u + 8ull * nsplit >= (unsigned long long)t.nodes_max || depth_max <= TOP_L) bad = 1u;
        if (bad) s_bad = 1u;
        __syncthreads();
        if (s_bad != 0u) {
            if (tid == 0 && blockIdx.x == 0) { t.top_flag[0] = 1u; hs[1] = 1u; hs[2] = 0u; hs[3] = 0u; __threadfence_system(); hs[0] = t.build_id; }
            return;
        }
    }
    // ---- levels: split nodes per level, first node id per level
    if (tid == 0) {
        uint32_t end = 1u;
        s_lvb[0] = 0u; s_lvb[1] = 1u;
        for (int d = 0; d <= TOP_L; d++) {
            const uint32_t internal = (exs[d == TOP_L ? TOP_COLS : top_off(d + 1)] - exs[top_off(d)]) & 0xFFFFu;
            s_int[d] = internal;
            end += 8u * internal;
            s_lvb[d + 2] = end;                                  // first id of level d + 2 = one past the last of level d + 1
        }
    }
    __syncthreads();
    // ---- every cell writes the records of its node (the CTAs of the grid - each has done the scan for itself - share the cells)
    for (uint32_t c = blockIdx.x * TN_THREADS + tid; c < TOP_COLS; c += gridDim.x * TN_THREADS) {
        const int d = top_level_of(c);
        const uint32_t off = top_off(d), q = c - off, cn = tot[c];
        const bool split = cn > sub_max;
        uint32_t id = 0u, pid = 0u;
        if (d > 0) {
            const uint32_t pc = top_off(d - 1) + (q >> 3);
            if (tot[pc] <= sub_max) { t.top_info[c] = NODE_END; continue; }          // the parent is not split: no node
            id = s_lvb[d] + 8u * ((exs[pc] - exs[top_off(d - 1)]) & 0xFFFFu) + (q & 7u);
            const uint32_t ppc = d > 1 ? top_off(d - 2) + (q >> 6) : 0u;
            pid = d > 1 ? s_lvb[d - 1] + 8u * ((exs[ppc] - exs[top_off(d - 2)]) & 0xFFFFu) + ((q >> 3) & 7u) : 0u;
        }
        t.top_info[c] = id | (split ? 0x80000000u : 0u);
        float lo[3], hi[3];
        top_cell(t.bounds, q, d, lo, hi);
        float4* cell = reinterpret_cast<float4*>(&t.rec2[id]);
        cell[0] = make_float4(lo[0], lo[1], lo[2], hi[0]);
        cell[1] = make_float4(hi[1], hi[2], __uint_as_float(pid), 0.f);
        t.rec1[id] = make_float4(0.f, 0.f, 0.f, 0.f);             // moments of empty nodes stay zero (tree.c:302-310)
        const uint32_t start = exb[c] - (uint32_t)d * M;          // every level holds all M in-box particles
        t.seg[id] = make_uint2(start, start + cn);
        if (d > 0) t.counts[(size_t)pid * 8u + (q & 7u)] = cn;
        // depth-first successor when the node is not opened: the next NON-EMPTY sibling, else the parent's successor
        uint32_t succ = NODE_END;
        {
            uint32_t qq = q;
            for (int lev = d; lev >= 1 && succ == NODE_END; lev--, qq >>= 3) {
                const uint32_t sib0 = top_off(lev) + (qq & ~7u);
                for (uint32_t o2 = 7u; o2 > (qq & 7u); o2--)
                    if (tot[sib0 + o2] > 0u) succ = s_lvb[lev] + 8u * ((exs[top_off(lev - 1) + (qq >> 3)] - exs[top_off(lev - 1)]) & 0xFFFFu) + o2;
            }
        }
        t.next[id] = succ;
        if (split) {
            t.child_base[id] = s_lvb[d + 1] + 8u * ((exs[c] - exs[off]) & 0xFFFFu);
            t.nmid[id] = make_float4(__fmul_rn(0.5f, __fadd_rn(lo[0], hi[0])), __fmul_rn(0.5f, __fadd_rn(lo[1], hi[1])),
                                     __fmul_rn(0.5f, __fadd_rn(lo[2], hi[2])), 0.f);
        } else {
            t.child_base[id] = NODE_END;
            if (sub_owns(t, cn)) t.sub_list[exs[c] >> 16] = make_uint2(id, (uint32_t)d);   // finished (moments included) by subtree_kernel
        }
        if (!(d == 0 && sub_owns(t, cn))) {                        // (the tree root as a subtree root is folded by subtree_kernel)
            if (!sub_owns(t, cn) && d <= XL_DEPTH_MAX && cn > t.xl_min) {   // very long fold: chunked exact evaluation (moments_xl)
                const uint32_t k = atomicAdd(&t.xl_count[d], 1u);
                if (k < t.xl_max) t.xl_nodes[(size_t)d * t.xl_max + k] = id;
            } else if (cn > ML_MIN && d > 0) t.ml_count[d] = 1u;
        }
        t.top_start[c] = start;
    }
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0) {
        // the per-build counters and level summaries, as CompactOp::total, root_kernel and open_level() leave them
        t.meta[M_NUM_INBOX] = M; t.meta[M_STATUS] = 0u; t.meta[M_DEPTH] = 0u; t.meta[M_FINAL] = 0u;
        t.meta[M_LV_BEGIN + 0] = 0u; t.meta[M_LV_END + 0] = 1u;
        for (int d = 0; d <= TOP_L; d++) {
            t.meta[M_LV_BEGIN + d + 1] = s_lvb[d + 1];
            t.meta[M_LV_END + d + 1] = s_lvb[d + 2];
            if (s_int[d] > 0u) t.meta[M_DEPTH] = (uint32_t)d + 1u; else t.meta[M_FINAL] = (uint32_t)d;
            volatile uint32_t* h = t.summary + 4 * d;
            h[1] = s_int[d]; h[2] = 0u; h[3] = s_lvb[d + 2];
        }
        t.meta[M_COUNT] = s_lvb[TOP_L + 2];
        *t.sub_count = exs[TOP_COLS] >> 16;                       // the subtree roots of the top levels (later levels append)
        hs[1] = 0u; hs[2] = 0u; hs[3] = s_lvb[TOP_L + 2];
        {   // what lets the HOST close the walk's root gate without a launch (tree_root_moments): the in-box bodies and the root's
            // filter radius^2 over its width^2, computed like finalize_kernel's (gravity.c:368-372)
            volatile uint32_t* hh = t.summary + 4 * (TOP_SUMMARY_SLOT + 1u);
            const float* b = t.bounds;
            const float ex = __fsub_rn(b[3], b[0]), ey = __fsub_rn(b[4], b[1]), ez = __fsub_rn(b[5], b[2]);
            const float slack = 4e-6f * (fabsf(b[0]) + fabsf(b[1]) + fabsf(b[2]) + fabsf(b[3]) + fabsf(b[4]) + fabsf(b[5]));
            const float diag = sqrtf(ex * ex + ey * ey + ez * ez);
            const float w = __fmul_rn((float)(1.0 / 3.0), __fadd_rn(__fadd_rn(ex, ey), ez));
            const float rin2 = (0.5f * diag + slack) * (0.5f * diag + slack) * 1.0001f + 1e-30f;
            hh[0] = M; hh[1] = __float_as_uint(rin2 / __fmul_rn(w, w)); hh[2] = t.top_flag[2];   // [2]: a body at the root's very centre
        }
        __threadfence_system();
        for (int d = 0; d <= TOP_L; d++) t.summary[4 * d] = t.build_id;
        hs[0] = t.build_id;
    }
}

// One CTA per range: the range's row of the prefix table, the cell starts and the "split" flags are folded into one
// shared-memory table (first slot of the cell for THIS range's particles), so a particle costs five shared-memory
// look-ups instead of a dozen scattered global ones.
__global__ void __launch_bounds__(TP_THREADS, 4) top_scatter_kernel(TreeDev t, const float4* __restrict__ pm4, uint32_t N, uint32_t range_len) {
    __shared__ uint32_t comb[TOP_COLS];
    if (t.top_flag[0] != 0u || t.meta[M_STATUS] != 0u) return;
    const uint32_t range = blockIdx.x;
    const uint32_t i_begin = range * range_len, i_end = min(N, i_begin + range_len);
    const uint32_t* __restrict__ rowp = t.top_table + (size_t)range * TOP_COLS;
    SlotRec* __restrict__ rec = t.srec[TOP_L];
    const bool need_owner = t.meta[M_LV_END + TOP_L + 1] != t.meta[M_LV_BEGIN + TOP_L + 1];   // level TOP_L has a cell to split
    for (uint32_t k = threadIdx.x; k < TOP_COLS; k += TP_THREADS) {
        const uint32_t inf = t.top_info[k];
        comb[k] = ((k == 0u ? 0u : t.top_start[k]) + rowp[k]) | ((inf != NODE_END && (inf >> 31)) ? 0x80000000u : 0u);
    }
    __syncthreads();
    constexpr int TS_BATCH = 4;                                   // particles a thread has in flight (their loads do not wait for each other)
    for (uint32_t ib = i_begin; ib < i_end; ib += TS_BATCH * TP_THREADS) {
    uint4 raws[TS_BATCH];
    float4 ps[TS_BATCH];
#pragma unroll
    for (int j = 0; j < TS_BATCH; j++) {
        const uint32_t i = ib + (uint32_t)j * TP_THREADS + threadIdx.x;
        raws[j] = i < i_end ? *reinterpret_cast<const uint4*>(&t.top_rank[i]) : make_uint4(0u, 0u, 0u, 0u);
        ps[j] = i < i_end ? pm4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < TS_BATCH; j++) {
        const uint32_t i = ib + (uint32_t)j * TP_THREADS + threadIdx.x;
        if (i >= i_end) continue;
        const uint4 raw = raws[j];
        const uint32_t code = raw.x & 0xFFFFu;
        const uint32_t c0 = comb[0];
        const uint32_t r0 = (c0 & 0x7FFFFFFFu) + (raw.x >> 16);
        if (code == TOP_NONE) { t.perm_out[i - r0] = i; continue; }
        const float4 p = ps[j];
        // every store is one whole 32-byte record; the folds of level d read the members of its nodes in srec[d]
        // (the level-0 order - the in-box particles in index order - is only read by the deferred fold of the root, which
        //  hardly ever runs: top_level0_kernel writes it then)
        const uint32_t c1 = comb[TOP_O1 + (code >> 9)], c2 = comb[TOP_O2 + (code >> 6)], c3 = comb[TOP_O3 + (code >> 3)], c4 = comb[TOP_O4 + code];
        const uint32_t p1 = (c1 & 0x7FFFFFFFu) + (raw.y & 0xFFFFu), p2 = (c2 & 0x7FFFFFFFu) + (raw.y >> 16);
        const uint32_t p3 = (c3 & 0x7FFFFFFFu) + (raw.z & 0xFFFFu), p4 = (c4 & 0x7FFFFFFFu) + (raw.z >> 16);
        // the particle descends while the cell it is in is split; digit d of the 48-bit path key sits at bit 45 - 3d
        uint32_t pos = r0;
        int d = 0;
        if (c0 >> 31) { pos = p1; d = 1; slot_store(&t.srec[1][p1], p, 0ull, i, NODE_END); }
        if (d == 1 && (c1 >> 31)) { pos = p2; d = 2; slot_store(&t.srec[2][p2], p, 0ull, i, NODE_END); }
        if (d == 2 && (c2 >> 31)) { pos = p3; d = 3; slot_store(&t.srec[3][p3], p, 0ull, i, NODE_END); }
        if (d == 3 && (c3 >> 31)) { pos = p4; d = 4; }
        const unsigned long long key = d == 0 ? 0ull : (unsigned long long)(code >> (3 * (TOP_L - d))) << (48 - 3 * d);
        uint32_t own = NODE_END;
        if (need_owner && d == TOP_L && (c4 >> 31)) own = t.top_info[TOP_O4 + code] & 0x7FFFFFFFu;
        slot_store(&rec[pos], p, key, i, own);
    }
    }
}

// The level-0 slot records of a build that took the one-pass top levels, written when the root's fold is needed after all
// (gate = {needed, done}): slot = in-box particles before this one, from the rank record and the prefix table of the build.
__global__ void __launch_bounds__(256) top_level0_kernel(TreeDev t, const float4* __restrict__ pm4, uint32_t N, uint32_t range_len,
                                                         const uint32_t* __restrict__ gate) {
    if (gate[0] == 0u || gate[1] != 0u) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const uint4 raw = *reinterpret_cast<const uint4*>(&t.top_rank[i]);
        if ((raw.x & 0xFFFFu) == TOP_NONE) continue;
        const uint32_t r0 = t.top_table[(size_t)(i / range_len) * TOP_COLS] + (raw.x >> 16);
        slot_store(&t.srec[0][r0], pm4[i], 0ull, i, NODE_END);
    }
}

}  // namespace sfe
