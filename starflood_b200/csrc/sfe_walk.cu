// sfe_walk.cu - Barnes-Hut tree walk (replaces solve_gravity_part_tree, gravity.c:245-663).
//
// The reference walks a parent-pointer tree with a per-particle path stack, visiting the children of an
// opened node in octant order 0..7 and taking every leaf (1..8 bodies) and every accepted internal node as
// one monopole.  Here the same depth-first visit order is produced without a stack: every
// node carries `next`, its successor when it is not opened (the next sibling, or the parent's successor
// after octant 7), built top-down by the tree build.  Since children are allocated 8 at a time the
// successor of child o < 7 is simply the next record.
//
// Targets are processed in Morton (slot) order, so the lanes of a warp hold neighbouring particles and
// follow nearly the same path: node records are fetched through L1 by all lanes at once and control flow
// stays mostly converged.  The 32-byte decision record {centre, width^2, next, child, rin^2, kind} is the only thing
// touched for an opened or empty node; the monopole {com, mass} is a second 16-byte record and the bounds
// (self-subtraction test, gravity.c:473-490) a third one that is read only when the target can be inside.
//
// Arithmetic:
//   * the opening decision reproduces gravity.c:353-377,448 exactly (no FMA contraction; the IEEE division
//     w^2/d^2 is only evaluated when the cheap comparison is within 1e-5 of the threshold), so the
//     interaction list of every particle is identical to the reference's;
//   * SFE_WALK_EXACT evaluates the monopole with IEEE div / sqrt and the reference's Kahan recurrence in the
//     same order: acc and pot are bit-identical to the reference;
//   * SFE_WALK_FAST uses rsqrt.approx and plain fp32 accumulation in the same visit order.
//
// Three kernels walk the same visits in the same order and produce the same bits (launch_walk() chooses):
//   walk_kernel     the records above in the build's breadth-first numbering; a rank's share of C2 on 2 / 4 GPUs
//   walk_p_kernel   ONE 32-byte record per visit in depth-first numbering (WalkRecP, made on demand): single-GPU solves
//   walk_sp_kernel  software-pipelined, for launches that fit the GPU in one wave (a rank's share on 8 GPUs, small N)
//
// Bound: on the SM.  EXACT: instruction issue (walk_p_kernel: 1.38e9 warp instructions per launch at N = 2^20, issue slots 82 %
// busy, 22 of 32 lanes active); FAST: the L1 data pipe (82 %: every lane fetches its own node).  The node table stays L1/L2
// resident, DRAM traffic is 0.4 % of the bytes of SURVEY 8(d)'s HBM model (32 B x node visits + particle I/O), which bench.py
// reports beside the binding roof.
#include "sfe_internal.cuh"
#include "sfe_tree.cuh"
#include "sfe_scan.cuh"
#include <stdlib.h>

namespace sfe {

constexpr int WALK_THREADS = 128;
constexpr int WALK_SP_CTAS = 7;     // resident CTAs per SM of walk_sp_kernel (72 registers)

struct KahanRef {   // gravity.c:771-799: y = F - com; t = sum + y; z = t - sum; com = z - y; sum = t
    float s, c;
    __device__ __forceinline__ void add(float x) {
        const float y = __fsub_rn(x, c);
        const float t = __fadd_rn(s, y);
        const float z = __fsub_rn(t, s);
        c = __fsub_rn(z, y);
        s = t;
    }
};

struct KahanRef2 {  // two of them in the lanes of packed fp32x2 instructions (same roundings, lane by lane)
    float2 s, c;
    __device__ __forceinline__ void add(float2 x) {
        const float2 y = __fadd2_rn(x, make_float2(-c.x, -c.y));
        const float2 t = __fadd2_rn(s, y);
        const float2 z = __fadd2_rn(t, make_float2(-s.x, -s.y));
        c = __fadd2_rn(z, make_float2(-y.x, -y.y));
        s = t;
    }};

// ---------------------------------------------------------------------------------------------------------
// Correctly rounded 1/q, sqrt(q) and 1/sqrt-then-reciprocal for the EXACT flavour.  The reference computes
// inv_r2 = 1/q and inv_r1 = 1/sqrtf(q) with IEEE division and square root (gravity.c:754-759); __fdiv_rn /
// __fsqrt_rn reproduce them but expand to ~10 instructions each plus a slow-path guard, and the EXACT walk is
// issue-bound.  For operands away from the denormal / overflow range the Newton / Markstein sequences below give
// the same bits in 5 + 7 + 5 instructions.  "The same bits" is not argued, it is checked: the fast path is only
// taken for q in [2^-24, 2^16) (the reference workloads have q = r^2 + 1e-6 < 50), and
// sfe_selftest_exact_math() compares the sequences against the intrinsics for EVERY float of that range on the
// device - 40 binades, 3.4e8 values (tests/test_gpu_walk.py::test_exact_math_exhaustive).
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_rn_seq(float q) {
    float r = rcp_approx(q);
    float e = __fmaf_rn(-q, r, 1.0f);
    r = __fmaf_rn(r, e, r);
    e = __fmaf_rn(-q, r, 1.0f);
    return __fmaf_rn(r, e, r);
}
// s = RN(sqrt(q)).  (Seeding 1/s with the MUFU.RSQ estimate instead of a MUFU.RCP of s saves one instruction but
// is NOT exact: the exhaustive check found 40 mismatching inputs in 3.4e8, so 1/s goes through rcp_rn_seq.)
__device__ __forceinline__ float sqrt_rn_seq(float q) {
    const float y = rsqrt_approx(q);
    float s = __fmul_rn(q, y);
    const float h = __fmul_rn(0.5f, y);
    float e = __fmaf_rn(-s, s, q);
    s = __fmaf_rn(e, h, s);
    e = __fmaf_rn(-s, s, q);
    return __fmaf_rn(e, h, s);
}
// the exhaustively verified operand range
constexpr float SEQ_LO = 5.9604644775390625e-08f;   // 2^-24
constexpr float SEQ_HI = 65536.0f;                   // 2^16
constexpr uint32_t SEQ_LO_BITS = 0x33800000u;        // bit pattern of 2^-24
constexpr uint32_t SEQ_BINADES = 40u;

__global__ void exact_math_selftest_kernel(unsigned long long* __restrict__ bad /* [3] */) {
    // every float in [SEQ_LO, SEQ_HI)
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < (SEQ_BINADES << 23); k += gridDim.x * blockDim.x) {
        const float q = __uint_as_float(SEQ_LO_BITS + k);
        const float s = sqrt_rn_seq(q);
        const float s_ref = __fsqrt_rn(q);
        if (__float_as_uint(rcp_rn_seq(q)) != __float_as_uint(__fdiv_rn(1.0f, q))) atomicAdd(&bad[0], 1ull);
        if (__float_as_uint(s) != __float_as_uint(s_ref)) atomicAdd(&bad[1], 1ull);
        if (__float_as_uint(rcp_rn_seq(s_ref)) != __float_as_uint(__fdiv_rn(1.0f, s_ref))) atomicAdd(&bad[2], 1ull);
    }
}

// theta <= (d2 > 0 ? fl(w2 / d2) : 0)      (gravity.c:377,448)
__device__ __forceinline__ bool opens(float w2, float d2, float theta) {
    if (!(0.0f < d2)) return theta <= 0.0f;
    const float rhs = theta * d2;
    if (w2 > rhs * 1.00001f) return true;
    if (w2 < rhs * 0.99999f) return false;
    return theta <= __fdiv_rn(w2, d2);
}

// Multi-rank routing: a rank walks a contiguous range of SLOTS - neighbours in Morton order, so its warps stay
// as coherent as on one GPU, which a slice of particle INDICES is not (indices are random in space) - and delivers every
// result straight into the window of the rank that OWNS the particle (index slice [r*N/world, (r+1)*N/world)), with
// plain stores over NVLink peer pointers; the arrival flags follow in peer_barrier_kernel.
struct WalkRoute { char* base[SFE_PEER_MAX]; uint64_t acc_off, pot_off; uint32_t world, N; };

__device__ __forceinline__ void load_rec0(uint32_t (&r)[8], const WalkRec0* p) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
// the cell of a node (rare path: self-subtraction test): one request for the whole 32-byte record
__device__ __forceinline__ void load_rec2(float4& lo, float4& hi, const WalkRec2* p) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w) : "l"(p));
}
__device__ __forceinline__ float4 load_rec1(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// RP_* of the tree the context holds, copied into constant memory behind every build (walk_publish_recp): warp-uniform values
// the loop reads as constant operands instead of keeping them in (or spilling them from) its 32 registers
__constant__ uint32_t c_recp_par[RP_WORDS];
int walk_publish_recp(Ctx* c, cudaStream_t s, const char* fn) {
    SFE_CUDA(fn, cudaMemcpyToSymbolAsync(c_recp_par, c->tree.recp_par.p, RP_WORDS * 4, 0, cudaMemcpyDeviceToDevice, s));
    return SFE_SUCCESS;
}

// the walk of one target on the breadth-first records {rec0, rec1, rec2}: walk_kernel's loop (also walk_p_kernel's fallback)
template <int FLAVOUR>
__device__ __forceinline__ void v1_loop(const TreeDev& t, const uint32_t M, const float x, const float y, const float z, const float mi,
                                        const float eps2, const float theta, float& fx_, float& fy_, float& fz_, float& fp_) {
    KahanRef2 kxy = {{0.f, 0.f}, {0.f, 0.f}}, kzp = {{0.f, 0.f}, {0.f, 0.f}};   // (acc x, acc y) and (acc z, pot)
    float fx = 0.f, fy = 0.f, fz = 0.f, fp = 0.f;

    const WalkRec0* __restrict__ rec0 = t.rec0;
    const float4* __restrict__ rec1 = t.rec1;
    const WalkRec2* __restrict__ rec2 = t.rec2;

    const bool seq_ok = eps2 >= SEQ_LO;
    uint32_t n = M == 0u ? NODE_END : 0u;   // an empty root is never evaluated (gravity.c:324-328)
    while (n != NODE_END) {
        // one 256-bit load = one L1 sector: {a[4], next, child, rin2, kind}
        uint32_t r[8];
        load_rec0(r, rec0 + n);
        const float a0 = __uint_as_float(r[0]), a1 = __uint_as_float(r[1]), a2 = __uint_as_float(r[2]), a3 = __uint_as_float(r[3]);
        const uint32_t next = r[4], kind = r[7];
        const float rin2 = __uint_as_float(r[6]);
        float rx, ry, rz, mj;
        bool maybe_inside;
        if (kind == KIND_INTERNAL) {
            const float2 oxy = __fadd2_rn(make_float2(a0, a1), make_float2(-x, -y));   // packed: same roundings, fewer issue slots
            const float2 osq = __fmul2_rn(oxy, oxy);
            const float oz = __fsub_rn(a2, z);
            const float d2 = __fadd_rn(__fadd_rn(osq.x, osq.y), __fmul_rn(oz, oz));
            if (opens(a3, d2, theta)) {                                          // gravity.c:448-453 (first non-empty child)
                n = r[5];
                continue;
            }
            const float4 cm = __ldg(&rec1[n]);                                   // accepted: monopole of the cell
            rx = cm.x; ry = cm.y; rz = cm.z; mj = cm.w;
            maybe_inside = d2 <= rin2;
        } else {                                                                 // leaf: the record IS the monopole
            rx = a0; ry = a1; rz = a2; mj = a3;
            const float ux = x - rx, uy = y - ry, uz = z - rz;
            maybe_inside = fmaxf(fmaxf(fabsf(ux), fabsf(uy)), fabsf(uz)) <= rin2;      // a cube about the COM (rin2 = -1 for a single body)
        }
        // monopole (gravity.c:463-548)
        if (maybe_inside) {
            float4 lo, hi;                                       // lo = {lo.x, lo.y, lo.z, hi.x}, hi = {hi.y, hi.z, -, -}
            load_rec2(lo, hi, &rec2[n]);
            if (lo.x <= x && x <= lo.w && lo.y <= y && y <= hi.x && lo.z <= z && z <= hi.y) {
                rx = __fsub_rn(__fmul_rn(rx, mj), __fmul_rn(mi, x));              // gravity.c:478-484
                ry = __fsub_rn(__fmul_rn(ry, mj), __fmul_rn(mi, y));
                rz = __fsub_rn(__fmul_rn(rz, mj), __fmul_rn(mi, z));
                mj = __fsub_rn(mj, mi);
                const float inv = 0.0f < mj ? __fdiv_rn(1.0f, mj) : 1.0f;
                rx = __fmul_rn(rx, inv); ry = __fmul_rn(ry, inv); rz = __fmul_rn(rz, inv);
            }
        }
        if (FLAVOUR == (int)SFE_WALK_EXACT) {
            // Packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2 round each lane to nearest like their scalar forms, so every
            // value below is the bit pattern of the scalar sequence): the EXACT walk is bound by issue slots, and a packed
            // instruction does two of its independent operations in one.
            const float2 dxy = __fadd2_rn(make_float2(x, y), make_float2(-rx, -ry));
            const float2 sq = __fmul2_rn(dxy, dxy);
            const float dz = __fsub_rn(z, rz);
            const float r2 = __fadd_rn(__fadd_rn(sq.x, sq.y), __fmul_rn(dz, dz));
            const float q = __fadd_rn(r2, eps2);
            float inv_r2, inv_r1;
            if (seq_ok && q < SEQ_HI) {      // q >= eps2 >= SEQ_LO here
                // rcp_rn_seq(q) and rcp_rn_seq(sqrt_rn_seq(q)) side by side in the two lanes
                const float sr = sqrt_rn_seq(q);
                const float2 nqs = make_float2(-q, -sr), one = make_float2(1.0f, 1.0f);
                float2 r = make_float2(rcp_approx(q), rcp_approx(sr));
                float2 e = __ffma2_rn(nqs, r, one);
                r = __ffma2_rn(r, e, r);
                e = __ffma2_rn(nqs, r, one);
                r = __ffma2_rn(r, e, r);
                inv_r2 = r.x; inv_r1 = r.y;
            } else {
                inv_r2 = __fdiv_rn(1.0f, q);
                inv_r1 = __fdiv_rn(1.0f, __fsqrt_rn(q));
            }
            const float pj = __fmul_rn(mj, inv_r1);
            const float w = __fmul_rn(pj, inv_r2);
            // (scalar products on purpose: ptxas contracts a mul.rn.f32x2 that feeds an add.rn.f32x2 into one FFMA2 -
            // one rounding instead of two - which it never does to the scalar .rn forms; seen in SASS, caught by the golden test)
            const float2 fxy = make_float2(__fmul_rn(w, dxy.x), __fmul_rn(w, dxy.y));
            const float2 fzp = make_float2(__fmul_rn(w, dz), pj);
            kxy.add(fxy); kzp.add(fzp);
        } else {
            const float dx = __fsub_rn(x, rx), dy = __fsub_rn(y, ry), dz = __fsub_rn(z, rz);
            const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float q = __fadd_rn(r2, eps2);
            const float inv = rsqrt_approx(q);
            const float pj = mj * inv;
            const float w = pj * (inv * inv);
            fx = fmaf(w, dx, fx); fy = fmaf(w, dy, fy); fz = fmaf(w, dz, fz);
            fp += pj;
        }
        n = next;
    }
    if (FLAVOUR == (int)SFE_WALK_EXACT) { fx = kxy.s.x; fy = kxy.s.y; fz = kzp.s.x; fp = kzp.s.y; }
    fx_ = fx; fy_ = fy; fz_ = fz; fp_ = fp;
}

template <int FLAVOUR, bool ROUTED>
__global__ void __launch_bounds__(WALK_THREADS, 16)
walk_kernel(TreeDev t, const float4* __restrict__ pm4, const uint32_t* __restrict__ slots /* null: the range from s0 */,
            uint32_t s0, uint32_t nslots, float G, float eps2, float theta,
            float* __restrict__ acc, float* __restrict__ pot, WalkRoute route) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nslots) return;
    if (t.meta[M_STATUS] != 0u) return;          // the build failed (tree.c:180-189): nothing consistent to walk; tree_status() raises it
    const uint32_t M = t.meta[M_NUM_INBOX];
    const uint32_t s = slots ? slots[k] : s0 + k;
    const uint32_t fin = t.meta[M_FINAL];
    uint32_t idx; float4 me;
    if (s < M) { const SlotRec r = slot_load(&t.srec[fin][s]); idx = r.idx; me = r.pm; }   // one sector: particle and index
    else { idx = t.perm_out[s - M]; me = pm4[idx]; }
    const float x = me.x, y = me.y, z = me.z, mi = me.w;

    float fx, fy, fz, fp;
    v1_loop<FLAVOUR>(t, M, x, y, z, mi, eps2, theta, fx, fy, fz, fp);
    const float gm = __fmul_rn(-G, mi);                                            // gravity.c:656-659
    if (ROUTED) {
        // owner r of particle idx: (N*r)/world <= idx < (N*(r+1))/world
        uint32_t r = (uint32_t)(((uint64_t)idx * route.world) / route.N);
        while (idx < (uint32_t)(((uint64_t)route.N * r) / route.world)) r--;
        while (idx >= (uint32_t)(((uint64_t)route.N * (r + 1u)) / route.world)) r++;
        acc = reinterpret_cast<float*>(route.base[r] + route.acc_off);
        pot = reinterpret_cast<float*>(route.base[r] + route.pot_off);
    }
    acc[3u * idx + 0u] = __fmul_rn(gm, fx);
    acc[3u * idx + 1u] = __fmul_rn(gm, fy);
    acc[3u * idx + 2u] = __fmul_rn(gm, fz);
    pot[idx] = __fmul_rn(-G, fp);
}

// ---------------------------------------------------------------------------------------------------------
// walk_p_kernel: walk_kernel on the pre-order records (WalkRecP, sfe_tree.cuh): ONE 32-byte load per visit.  An accepted cell
// no longer fetches its COM record (15.9 M of walk_kernel's 34.4 M scattered requests per launch at C2), the first child of an
// opened cell is the next record, and the cell's width^2 comes from its depth (root width^2 / 4^depth, verified per node by
// the build to RECP_TOL); an opening test closer than RECP_BAND to theta, a cell the build could not verify, and the
// self-subtraction test go to the breadth-first records through bfs_of[].  Same visits, same arithmetic, same bits.
// When the build could not number the tree (RP_BAD: a degenerate root box) the kernel runs walk_kernel's loop instead.
// RANGE_OK: every q = r^2 + eps^2 this thread can meet lies in the verified range of the Newton sequences (decided once per
// warp from the root box, as in walk_sp_kernel), so the loop carries no IEEE fallback.
template <int FLAVOUR, bool RANGE_OK>
__device__ __forceinline__ void p_loop(const TreeDev& t, uint32_t n, const float x, const float y, const float z, const float mi,
                                       const float eps2, const float theta, float& fx, float& fy, float& fz, float& fp) {
    KahanRef2 kxy = {{0.f, 0.f}, {0.f, 0.f}}, kzp = {{0.f, 0.f}, {0.f, 0.f}};   // (acc x, acc y) and (acc z, pot)
    fx = fy = fz = fp = 0.f;
    const WalkRecP* __restrict__ recp = t.recp;
    while (n != RECP_END) {
        uint32_t r[8];
        load_rec0(r, reinterpret_cast<const WalkRec0*>(recp + n));     // one 256-bit load: the whole visit
        const uint32_t w = r[7];
        const float rx0 = __uint_as_float(r[0]), ry0 = __uint_as_float(r[1]), rz0 = __uint_as_float(r[2]), mj0 = __uint_as_float(r[3]);
        const uint32_t here = n;
        bool maybe_inside = false, eval = true;
        uint32_t succ = w & RECP_END;
        if ((int)w < 0) {                                                // a cell: {COM, mass, centre}
            const float2 oxy = __fadd2_rn(make_float2(__uint_as_float(r[4]), __uint_as_float(r[5])), make_float2(-x, -y));
            const float2 osq = __fmul2_rn(oxy, oxy);
            const float oz = __fsub_rn(__uint_as_float(r[6]), z);
            const float d2 = __fadd_rn(__fadd_rn(osq.x, osq.y), __fmul_rn(oz, oz));
            float w2 = __uint_as_float(c_recp_par[RP_W0] - (w & 0x1F000000u));   // root width^2 / 4^depth
            float rin2 = __uint_as_float(c_recp_par[RP_KR]) * w2;           // (the filter of the self-subtraction: a multiple of width^2)
            // theta <= fl(w^2 / d^2) (gravity.c:377,448): the sign of w^2 - theta d^2 settles it unless that is within RECP_BAND
            const float rhs = theta * d2, diff = w2 - rhs;
            bool open = diff > 0.0f;
            if (!(w >= 0xC0000000u && fabsf(diff) > rhs * RECP_BAND && 0.0f < d2)) {        // rare: the exact width^2
                const WalkRec0* e = t.rec0 + t.bfs_of[n];
                w2 = __ldg(&e->a[3]); rin2 = __ldg(&e->rin2);
                open = opens(w2, d2, theta);
            }
            // an opened cell goes on with its first non-empty child, the next record (gravity.c:448-453) - a select, not a
            // `continue`: the divergent exit out of two nested regions cost ~16 % of the loop's instructions in bookkeeping
            eval = !open;
            succ = open ? n + 1u : succ;
            maybe_inside = d2 <= rin2;
        } else {                                                         // a leaf: the record IS the monopole
            const float ux = x - rx0, uy = y - ry0, uz = z - rz0;
            maybe_inside = fmaxf(fmaxf(fabsf(ux), fabsf(uy)), fabsf(uz)) <= __uint_as_float(r[4]);   // (-1 for a single body)
        }
        n = succ;
        if (!eval) continue;
        float rx = rx0, ry = ry0, rz = rz0, mj = mj0;
        // monopole (gravity.c:463-548)
        if (maybe_inside) {
            float4 lo, hi;                                       // lo = {lo.x, lo.y, lo.z, hi.x}, hi = {hi.y, hi.z, -, -}
            load_rec2(lo, hi, &t.rec2[t.bfs_of[here]]);
            if (lo.x <= x && x <= lo.w && lo.y <= y && y <= hi.x && lo.z <= z && z <= hi.y) {
                rx = __fsub_rn(__fmul_rn(rx, mj), __fmul_rn(mi, x));              // gravity.c:478-484
                ry = __fsub_rn(__fmul_rn(ry, mj), __fmul_rn(mi, y));
                rz = __fsub_rn(__fmul_rn(rz, mj), __fmul_rn(mi, z));
                mj = __fsub_rn(mj, mi);
                const float inv = 0.0f < mj ? __fdiv_rn(1.0f, mj) : 1.0f;
                rx = __fmul_rn(rx, inv); ry = __fmul_rn(ry, inv); rz = __fmul_rn(rz, inv);
            }
        }
        if (FLAVOUR == (int)SFE_WALK_EXACT) {
            const float2 dxy = __fadd2_rn(make_float2(x, y), make_float2(-rx, -ry));
            const float2 sq = __fmul2_rn(dxy, dxy);
            const float dz = __fsub_rn(z, rz);
            const float r2 = __fadd_rn(__fadd_rn(sq.x, sq.y), __fmul_rn(dz, dz));
            const float q = __fadd_rn(r2, eps2);
            float inv_r2, inv_r1;
            if (RANGE_OK || q < SEQ_HI) {    // (q >= eps2 >= SEQ_LO: the caller's business)
                const float sr = sqrt_rn_seq(q);
                const float2 nqs = make_float2(-q, -sr), one = make_float2(1.0f, 1.0f);
                float2 rr = make_float2(rcp_approx(q), rcp_approx(sr));
                float2 e = __ffma2_rn(nqs, rr, one);
                rr = __ffma2_rn(rr, e, rr);
                e = __ffma2_rn(nqs, rr, one);
                rr = __ffma2_rn(rr, e, rr);
                inv_r2 = rr.x; inv_r1 = rr.y;
            } else {
                inv_r2 = __fdiv_rn(1.0f, q);
                inv_r1 = __fdiv_rn(1.0f, __fsqrt_rn(q));
            }
            const float pj = __fmul_rn(mj, inv_r1);
            const float wgt = __fmul_rn(pj, inv_r2);
            // (scalar products on purpose: see walk_kernel)
            const float2 fxy = make_float2(__fmul_rn(wgt, dxy.x), __fmul_rn(wgt, dxy.y));
            const float2 fzp = make_float2(__fmul_rn(wgt, dz), pj);
            kxy.add(fxy); kzp.add(fzp);
        } else {
            const float dx = __fsub_rn(x, rx), dy = __fsub_rn(y, ry), dz = __fsub_rn(z, rz);
            const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float q = __fadd_rn(r2, eps2);
            const float inv = rsqrt_approx(q);
            const float pj = mj * inv;
            const float wgt = pj * (inv * inv);
            fx = fmaf(wgt, dx, fx); fy = fmaf(wgt, dy, fy); fz = fmaf(wgt, dz, fz);
            fp += pj;
        }
    }
    if (FLAVOUR == (int)SFE_WALK_EXACT) { fx = kxy.s.x; fy = kxy.s.y; fz = kzp.s.x; fp = kzp.s.y; }
}

template <int FLAVOUR, bool ROUTED>
__global__ void __launch_bounds__(WALK_THREADS, 16)
walk_p_kernel(TreeDev t, const float4* __restrict__ pm4, const uint32_t* __restrict__ slots /* null: the range from s0 */,
              uint32_t s0, uint32_t nslots, float G, float eps2, float theta,
              float* __restrict__ acc, float* __restrict__ pot, WalkRoute route) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nslots) return;
    if (t.meta[M_STATUS] != 0u) return;
    const uint32_t M = t.meta[M_NUM_INBOX];
    const uint32_t s = slots ? slots[k] : s0 + k;
    const uint32_t fin = t.meta[M_FINAL];
    uint32_t idx; float4 me;
    if (s < M) { const SlotRec r = slot_load(&t.srec[fin][s]); idx = r.idx; me = r.pm; }   // one sector: particle and index
    else { idx = t.perm_out[s - M]; me = pm4[idx]; }
    const float x = me.x, y = me.y, z = me.z, mi = me.w;
    const uint32_t n0 = M == 0u ? RECP_END : 0u;   // an empty root is never evaluated (gravity.c:324-328)
    // every source (a COM, with or without the target taken out) lies inside the root box: the farthest corner bounds r^2
    bool range_ok = true;
    if (FLAVOUR == (int)SFE_WALK_EXACT) {
        const float ex = fmaxf(fabsf(x - t.bounds[0]), fabsf(x - t.bounds[3]));
        const float ey = fmaxf(fabsf(y - t.bounds[1]), fabsf(y - t.bounds[4]));
        const float ez = fmaxf(fabsf(z - t.bounds[2]), fabsf(z - t.bounds[5]));
        range_ok = eps2 >= SEQ_LO && (ex * ex + ey * ey + ez * ez) * 1.01f + eps2 < 0.5f * SEQ_HI;
    }
    float fx, fy, fz, fp;
    if (c_recp_par[RP_BAD] != 0u) v1_loop<FLAVOUR>(t, M, x, y, z, mi, eps2, theta, fx, fy, fz, fp);   // the build could not number this tree
    else if (__all_sync(__activemask(), range_ok)) p_loop<FLAVOUR, true>(t, n0, x, y, z, mi, eps2, theta, fx, fy, fz, fp);
    else p_loop<FLAVOUR, false>(t, n0, x, y, z, mi, eps2, theta, fx, fy, fz, fp);
    const float gm = __fmul_rn(-G, mi);                                            // gravity.c:656-659
    if (ROUTED) {
        uint32_t rk = (uint32_t)(((uint64_t)idx * route.world) / route.N);
        while (idx < (uint32_t)(((uint64_t)route.N * rk) / route.world)) rk--;
        while (idx >= (uint32_t)(((uint64_t)route.N * (rk + 1u)) / route.world)) rk++;
        acc = reinterpret_cast<float*>(route.base[rk] + route.acc_off);
        pot = reinterpret_cast<float*>(route.base[rk] + route.pot_off);
    }
    acc[3u * idx + 0u] = __fmul_rn(gm, fx);
    acc[3u * idx + 1u] = __fmul_rn(gm, fy);
    acc[3u * idx + 2u] = __fmul_rn(gm, fz);
    pot[idx] = __fmul_rn(-G, fp);
}

// ---------------------------------------------------------------------------------------------------------
// walk_sp_kernel: the same walk as a three-stage software pipeline inside every thread, for launches whose duration is the
// length of ONE warp's chain of dependent visits (a rank's 1/G share of the targets: 7 warps per scheduler at G = 8, a lone
// warp needs ~630 cycles per visit in walk_kernel - fixed-latency dependencies 235, node loads 217, divergent branches
// 122; profiles/r02_walk_slice_ncu.txt).  Per loop iteration a thread
//     A   visits node n: opening decision, successor, and - when the node is accepted - its monopole as the SOURCE of B1;
//     B1  starts the source accepted by the previous visit: distance, q = r^2 + eps^2, sqrt(q), reciprocal seeds;
//     B2  finishes the source of two visits ago: Newton steps, force terms, Kahan accumulation;
// three independent dependency chains in one basic block (the opened / accepted / leaf cases are selects and predicated
// loads, not branches; only the rare cases - self-subtraction, an opening test within 1e-5 of theta, q outside the verified
// range - branch), so the scheduler fills the latency of each with the other two and the lanes of a warp stay converged.
// The record of `next` is requested as soon as it is known; an opened cell replaces the request by its first child's.
// Sources reach the accumulators in visit order through the pipeline, with the arithmetic of walk_kernel: same bits.
__device__ __forceinline__ void load_rec0_if(uint32_t (&r)[8], const WalkRec0* p, bool c) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %9, 0;\n\t@q ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t}"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 : "l"(p), "r"((uint32_t)c));
}
__device__ __forceinline__ void load_rec1_if(float& a, float& b, float& c_, float& d, const float4* p, bool c) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "+f"(a), "+f"(b), "+f"(c_), "+f"(d) : "l"(p), "r"((uint32_t)c));
}

struct SpOut { float fx, fy, fz, fp; };

// RANGE_OK: every q = r^2 + eps^2 this thread can meet lies in the verified range of the Newton sequences (decided once per
// warp from the root box), so the loop carries no IEEE fallback.
template <int FLAVOUR, bool RANGE_OK>
__device__ __forceinline__ SpOut sp_loop(const WalkRec0* __restrict__ rec0, const float4* __restrict__ rec1,
                                         const WalkRec2* __restrict__ rec2, uint32_t n, float x, float y, float z, float mi,
                                         float eps2, float theta) {
    KahanRef2 kxy = {{0.f, 0.f}, {0.f, 0.f}}, kzp = {{0.f, 0.f}, {0.f, 0.f}};   // (acc x, acc y) and (acc z, pot)
    float fx = 0.f, fy = 0.f, fz = 0.f, fp = 0.f;
    uint32_t ra[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, rb[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    load_rec0_if(ra, rec0, n != NODE_END);
    float sx = 0.f, sy = 0.f, sz = 0.f, sm = 0.f; bool v1 = false;                     // A  -> B1
    float2 b_d = make_float2(0.f, 0.f), b_r = make_float2(1.f, 1.f);                   // B1 -> B2
    float b_dz = 0.f, b_q = 1.f, b_sr = 1.f, b_m = 0.f; bool v2 = false, b_ieee = false;

    auto step = [&](uint32_t (&cur)[8], uint32_t (&nxt)[8]) {
        // ---- A, loads: the successor's record is wanted by the next iteration whatever this one decides (87 % of the
        //      visits)
        const bool active = n != NODE_END;
        const uint32_t next = cur[4], child = cur[5], kind = cur[7];
        const uint32_t succ = (active && next != NODE_END) ? next : 0u;
        load_rec0(nxt, rec0 + succ);
        // ---- B2: gravity.c:754-799 for the source of two visits ago
        if (FLAVOUR == (int)SFE_WALK_EXACT) {
            const float2 nqs = make_float2(-b_q, -b_sr), one = make_float2(1.0f, 1.0f);
            float2 rr = b_r;
            float2 e = __ffma2_rn(nqs, rr, one);
            rr = __ffma2_rn(rr, e, rr);
            e = __ffma2_rn(nqs, rr, one);
            rr = __ffma2_rn(rr, e, rr);
            const float inv_r2 = (!RANGE_OK && b_ieee) ? b_r.x : rr.x, inv_r1 = (!RANGE_OK && b_ieee) ? b_r.y : rr.y;
            const float pj = __fmul_rn(b_m, inv_r1);
            const float w = __fmul_rn(pj, inv_r2);
            const float2 fxy = make_float2(__fmul_rn(w, b_d.x), __fmul_rn(w, b_d.y));   // scalar products: see walk_kernel
            const float2 fzp = make_float2(__fmul_rn(w, b_dz), pj);
            if (v2) { kxy.add(fxy); kzp.add(fzp); }
        } else {
            const float inv = b_r.x;
            const float pj = b_m * inv;
            const float w = pj * (inv * inv);
            if (v2) { fx = fmaf(w, b_d.x, fx); fy = fmaf(w, b_d.y, fy); fz = fmaf(w, b_dz, fz); fp += pj; }
        }
        // ---- B1: the source accepted by the previous visit
        bool odd_q = false;
        if (FLAVOUR == (int)SFE_WALK_EXACT) {
            const float2 dxy = __fadd2_rn(make_float2(x, y), make_float2(-sx, -sy));
            const float2 sq = __fmul2_rn(dxy, dxy);
            const float dz = __fsub_rn(z, sz);
            const float q = __fadd_rn(__fadd_rn(__fadd_rn(sq.x, sq.y), __fmul_rn(dz, dz)), eps2);
            const float sr = sqrt_rn_seq(q);
            b_d = dxy; b_dz = dz; b_q = q; b_sr = sr; b_r = make_float2(rcp_approx(q), rcp_approx(sr));
            if (!RANGE_OK) odd_q = v1 && !(eps2 >= SEQ_LO && q < SEQ_HI);   // IEEE forms below
        } else {
            const float dx = __fsub_rn(x, sx), dy = __fsub_rn(y, sy), dz = __fsub_rn(z, sz);
            const float q = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)), eps2);
            b_d = make_float2(dx, dy); b_dz = dz; b_r.x = rsqrt_approx(q);
        }
        b_m = sm; v2 = v1; b_ieee = false;
        // ---- A: opening decision (gravity.c:353-377,448): theta <= fl(w^2 / d^2), settled by the sign of w^2 - theta d^2
        //      unless that is within 1e-5 of zero (or d^2 is not positive) - then the IEEE division decides, below
        const float a0 = __uint_as_float(cur[0]), a1 = __uint_as_float(cur[1]), a2 = __uint_as_float(cur[2]), a3 = __uint_as_float(cur[3]);
        const float rin2 = __uint_as_float(cur[6]);
        const float2 oxy = __fadd2_rn(make_float2(a0, a1), make_float2(-x, -y));
        const float2 osq = __fmul2_rn(oxy, oxy);
        const float oz = __fsub_rn(a2, z);
        const float d2 = __fadd_rn(__fadd_rn(osq.x, osq.y), __fmul_rn(oz, oz));
        const bool cell = active && kind == KIND_INTERNAL;
        const float rhs = theta * d2;
        const float diff = a3 - rhs;
        const bool open_fast = cell && diff > 0.0f;
        const bool amb = cell && !(fabsf(diff) > rhs * 1.0e-5f && 0.0f < d2);
        load_rec0_if(nxt, rec0 + child, open_fast);      // an opened cell continues at its first non-empty child (gravity.c:448-453)
        float4 cm = make_float4(a0, a1, a2, a3);          // a leaf's record IS its monopole (gravity.c:331) ...
        load_rec1_if(cm.x, cm.y, cm.z, cm.w, rec1 + n, cell && !open_fast);   // ... an accepted cell's is the second record
        bool open = open_fast;
        // the target may be one of the node's bodies (rin2 < 0: single body)
        const bool near = active && (kind == KIND_INTERNAL ? d2 : fmaxf(fmaxf(fabsf(oxy.x), fabsf(oxy.y)), fabsf(oz))) <= rin2;
        // ---- the rare cases, all behind one branch
        if (amb || near || odd_q) {
            if (amb) {
                open = opens(a3, d2, theta);
                if (open != open_fast) { load_rec0(nxt, rec0 + (open ? child : succ)); if (!open) cm = load_rec1(rec1 + n); }
            }
            if (near && !open) {
                float4 lo, hi;
                load_rec2(lo, hi, &rec2[n]);
                if (lo.x <= x && x <= lo.w && lo.y <= y && y <= hi.x && lo.z <= z && z <= hi.y) {
                    cm.x = __fsub_rn(__fmul_rn(cm.x, cm.w), __fmul_rn(mi, x));            // gravity.c:478-484
                    cm.y = __fsub_rn(__fmul_rn(cm.y, cm.w), __fmul_rn(mi, y));
                    cm.z = __fsub_rn(__fmul_rn(cm.z, cm.w), __fmul_rn(mi, z));
                    cm.w = __fsub_rn(cm.w, mi);
                    const float inv = 0.0f < cm.w ? __fdiv_rn(1.0f, cm.w) : 1.0f;
                    cm.x = __fmul_rn(cm.x, inv); cm.y = __fmul_rn(cm.y, inv); cm.z = __fmul_rn(cm.z, inv);
                }
            }
            if (!RANGE_OK && odd_q) { b_r = make_float2(__fdiv_rn(1.0f, b_q), __fdiv_rn(1.0f, __fsqrt_rn(b_q))); b_ieee = true; }
        }
        sx = cm.x; sy = cm.y; sz = cm.z; sm = cm.w; v1 = active && !open;
        n = open ? child : (active ? next : NODE_END);
    };
    while (true) {
        if (!(n != NODE_END || v1 || v2)) break;
        step(ra, rb);
        if (!(n != NODE_END || v1 || v2)) break;
        step(rb, ra);
    }
    if (FLAVOUR == (int)SFE_WALK_EXACT) { fx = kxy.s.x; fy = kxy.s.y; fz = kzp.s.x; fp = kzp.s.y; }
    return SpOut{fx, fy, fz, fp};
}

template <int FLAVOUR, bool ROUTED>
__global__ void __launch_bounds__(WALK_THREADS, WALK_SP_CTAS)
walk_sp_kernel(TreeDev t, const float4* __restrict__ pm4, const uint32_t* __restrict__ slots /* null: the range from s0 */,
               uint32_t s0, uint32_t nslots, float G, float eps2, float theta,
               float* __restrict__ acc, float* __restrict__ pot, WalkRoute route) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nslots) return;
    if (t.meta[M_STATUS] != 0u) return;
    const uint32_t M = t.meta[M_NUM_INBOX];
    const uint32_t s = slots ? slots[k] : s0 + k;
    const uint32_t fin = t.meta[M_FINAL];
    uint32_t idx; float4 me;
    if (s < M) { const SlotRec r = slot_load(&t.srec[fin][s]); idx = r.idx; me = r.pm; }
    else { idx = t.perm_out[s - M]; me = pm4[idx]; }
    const float x = me.x, y = me.y, z = me.z, mi = me.w;
    const uint32_t n0 = M == 0u ? NODE_END : 0u;   // an empty root is never evaluated (gravity.c:324-328)

    // every source (a COM, with or without the target taken out) lies inside the root box: the farthest corner bounds r^2
    bool range_ok = false;
    if (FLAVOUR == (int)SFE_WALK_EXACT) {
        const float ex = fmaxf(fabsf(x - t.bounds[0]), fabsf(x - t.bounds[3]));
        const float ey = fmaxf(fabsf(y - t.bounds[1]), fabsf(y - t.bounds[4]));
        const float ez = fmaxf(fabsf(z - t.bounds[2]), fabsf(z - t.bounds[5]));
        range_ok = eps2 >= SEQ_LO && (ex * ex + ey * ey + ez * ez) * 1.01f + eps2 < 0.5f * SEQ_HI;
    }
    SpOut o;
    if (FLAVOUR != (int)SFE_WALK_EXACT || __all_sync(__activemask(), range_ok))
        o = sp_loop<FLAVOUR, true>(t.rec0, t.rec1, t.rec2, n0, x, y, z, mi, eps2, theta);
    else
        o = sp_loop<FLAVOUR, false>(t.rec0, t.rec1, t.rec2, n0, x, y, z, mi, eps2, theta);
    const float gm = __fmul_rn(-G, mi);                                            // gravity.c:656-659
    if (ROUTED) {
        uint32_t r = (uint32_t)(((uint64_t)idx * route.world) / route.N);
        while (idx < (uint32_t)(((uint64_t)route.N * r) / route.world)) r--;
        while (idx >= (uint32_t)(((uint64_t)route.N * (r + 1u)) / route.world)) r++;
        acc = reinterpret_cast<float*>(route.base[r] + route.acc_off);
        pot = reinterpret_cast<float*>(route.base[r] + route.pot_off);
    }
    acc[3u * idx + 0u] = __fmul_rn(gm, o.fx);
    acc[3u * idx + 1u] = __fmul_rn(gm, o.fy);
    acc[3u * idx + 2u] = __fmul_rn(gm, o.fz);
    pot[idx] = __fmul_rn(-G, o.fp);
}

// Which kernel: walk_sp_kernel when all its CTAs are resident at once (then the launch lasts as long as one warp's chain of
// visits, which the pipeline shortens), walk_kernel otherwise (fewer loads and registers: more throughput).  Measured at
// N = 2^20 (profiles/r02_walk_sp.txt): N/16 targets 0.36 -> 0.24 ms, N/8 0.40 -> 0.34, N/4 0.60 -> 0.60, all 1.83 -> 1.95.
// SFE_WALK_SP=0 / 1 or sfe_walk_select() forces one of them (test / measurement hook).
static int g_walk_select = -2;     // -2: not read yet, -1: by size and flavour, 0 / 1 / 2: forced (SFE_WALK_SP or sfe_walk_select())
static bool walk_use_sp(Ctx* c, uint32_t ntargets) {
    if (g_walk_select == -2) { const char* e = getenv("SFE_WALK_SP"); g_walk_select = e ? atoi(e) : -1; }
    if (g_walk_select >= 0) return g_walk_select == 1;
    return (uint64_t)ntargets <= (uint64_t)c->sm_count * WALK_SP_CTAS * WALK_THREADS;
}

template <bool ROUTED>
static int launch_walk(Ctx* c, uint32_t flavour, uint32_t grid, cudaStream_t s, const TreeDev& dv, const float4* pm4,
                        const uint32_t* slots, uint32_t s0, uint32_t ns, float G, float eps2, float theta, float* acc, float* pot,
                        const WalkRoute& rt) {
    const bool sp = walk_use_sp(c, ns);
    // walk_p_kernel (one record per visit, depth-first numbering) where it pays for its records, which the first walk that
    // wants them makes: five launches over all slots and nodes, ~45 us per build at C2 (~30 us + 12 us per million particles).
    // [B200] C2, all targets: FAST 1.50 -> 1.30 ms (bound by the L1 data pipe: a third fewer requests), EXACT 1.68 -> 1.57 ms
    // (bound by issue slots: shorter paths) - 0.22 / 0.10 ns per target.  Used when that gain is worth 1.3 x the records:
    // single-GPU solves, FAST shares down to a quarter of C2, EXACT shares of C5 on 2 / 4 GPUs; the rest stays on walk_kernel /
    // walk_sp_kernel.  sfe_walk_select(2) forces it (tests).  For a tree the numbering cannot describe (a device flag:
    // degenerate root box) the kernel runs walk_kernel's loop instead.
    const double records_us = 30.0 + 12.0e-6 * (double)c->tree.N;
    const double gain_us = (double)ns * (flavour == SFE_WALK_FAST ? 2.2e-4 : 1.04e-4);
    const bool p_pays = c->tree.recp_ready || gain_us > 1.3 * records_us;          // (records that exist cost nothing more)
    if (!sp && (g_walk_select == 2 || (g_walk_select == -1 && p_pays)) && c->tree.nodes_max < RECP_END) {
        if (tree_ensure_recp(c, s, "sfe_bh")) return SFE_FAILURE;
        if (flavour == SFE_WALK_EXACT) walk_p_kernel<(int)SFE_WALK_EXACT, ROUTED><<<grid, WALK_THREADS, 0, s>>>(dv, pm4, slots, s0, ns, G, eps2, theta, acc, pot, rt);
        else walk_p_kernel<(int)SFE_WALK_FAST, ROUTED><<<grid, WALK_THREADS, 0, s>>>(dv, pm4, slots, s0, ns, G, eps2, theta, acc, pot, rt);
        return SFE_SUCCESS;
    }
#define SFE_WALK_GO(K, FL) K<(int)FL, ROUTED><<<grid, WALK_THREADS, 0, s>>>(dv, pm4, slots, s0, ns, G, eps2, theta, acc, pot, rt)
    if (flavour == SFE_WALK_EXACT) { if (sp) SFE_WALK_GO(walk_sp_kernel, SFE_WALK_EXACT); else SFE_WALK_GO(walk_kernel, SFE_WALK_EXACT); }
    else { if (sp) SFE_WALK_GO(walk_sp_kernel, SFE_WALK_FAST); else SFE_WALK_GO(walk_kernel, SFE_WALK_FAST); }
#undef SFE_WALK_GO
    return SFE_SUCCESS;
}

// slots whose particle index lies in [i0, i0+ni), in slot order (multi-GPU target slices)
struct SliceOp {
    TreeDev t; uint32_t N, i0, i1; uint32_t* out; uint32_t* count;
    __device__ uint32_t n() const { return N; }
    __device__ uint32_t idx_of(uint32_t s) const {
        const uint32_t M = t.meta[M_NUM_INBOX];
        return s < M ? t.srec[t.meta[M_FINAL]][s].idx : t.perm_out[s - M];
    }
    __device__ uint32_t load(uint32_t s) const { const uint32_t i = idx_of(s); return (i >= i0 && i < i1) ? 1u : 0u; }
    __device__ void store(uint32_t s, uint32_t ex) const { const uint32_t i = idx_of(s); if (i >= i0 && i < i1) out[ex] = s; }
    __device__ void total(uint32_t c) const { *count = c; }
};

// compacted slot list of a target slice; lives in the tree's rank buffer (free once the build is done)
static int walk_slice_slots(Ctx* c, uint32_t N, uint32_t i0, uint32_t ni, cudaStream_t s, uint32_t** slots_out) {
    Tree& t = c->tree;
    if (t.slice_count.reserve(4, "sfe_bh")) return SFE_FAILURE;
    SliceOp op{tree_dev(c), N, i0, i0 + ni, t.rank.as<uint32_t>(), t.slice_count.as<uint32_t>()};
    if (run_scan(c, op, N, s, "sfe_bh")) return SFE_FAILURE;
    *slots_out = t.rank.as<uint32_t>();
    return SFE_SUCCESS;
}

int launch_bh(Ctx* c, const float4* pm4, uint32_t N, uint32_t i0, uint32_t ni, float* acc, float* pot,
              float G, float eps2, float theta, uint32_t flavour, cudaStream_t s) {
    const char* fn = "sfe_bh";
    Tree& t = c->tree;
    if (!t.valid || t.N != N) { set_error(fn, "no tree for N=%u: call sfe_tree_build() first", N); return SFE_FAILURE; }
    if ((uint64_t)i0 + ni > N) { set_error(fn, "target slice [%u,%u) exceeds N=%u", i0, i0 + ni, N); return SFE_FAILURE; }
    if (flavour != SFE_WALK_EXACT && flavour != SFE_WALK_FAST) { set_error(fn, "unknown walk flavour %u", flavour); return SFE_FAILURE; }
    if (ni == 0) return SFE_SUCCESS;
    if (tree_root_moments(c, pm4, i0, ni, theta, false, s)) return SFE_FAILURE;
    TreeDev dv = tree_dev(c);
    uint32_t* slots = nullptr;
    if (ni != N && walk_slice_slots(c, N, i0, ni, s, &slots)) return SFE_FAILURE;
    const uint32_t grid = (ni + WALK_THREADS - 1) / WALK_THREADS;
    WalkRoute none; none.world = 0u; none.N = N; none.acc_off = none.pot_off = 0u;
    if (launch_walk<false>(c, flavour, grid, s, dv, pm4, slots, 0u, ni, G, eps2, theta, acc, pot, none)) return SFE_FAILURE;
    return check_launch(fn);
}

// Collective over the open peer window: this rank walks the slots [rank*N/world, (rank+1)*N/world) of the (identical)
// tree every rank has built and stores acc / pot of each particle into its owner's window at acc_off + 12*i, pot_off + 4*i;
// after the flag exchange every rank holds acc / pot of its own index slice.
int launch_bh_routed(Ctx* c, const float4* pm4, uint32_t N, uint64_t acc_off, uint64_t pot_off,
                     float G, float eps2, float theta, uint32_t flavour, cudaStream_t s) {
    const char* fn = "sfe_dev_bh_peer";
    Tree& t = c->tree;
    PeerState& pr = c->peer;
    if (!pr.open) { set_error(fn, "the peer window is not open: call sfe_peer_create() and sfe_peer_open() first"); return SFE_FAILURE; }
    if (!t.valid || t.N != N) { set_error(fn, "no tree for N=%u: call sfe_tree_build() first", N); return SFE_FAILURE; }
    if (flavour != SFE_WALK_EXACT && flavour != SFE_WALK_FAST) { set_error(fn, "unknown walk flavour %u", flavour); return SFE_FAILURE; }
    if ((acc_off & 3u) || (pot_off & 3u) || acc_off + 12ull * N > pr.bytes || pot_off + 4ull * N > pr.bytes) {
        set_error(fn, "acc / pot ranges [%llu,+%llu) [%llu,+%llu) do not fit the %llu-byte window", (unsigned long long)acc_off,
                  12ull * N, (unsigned long long)pot_off, 4ull * N, (unsigned long long)pr.bytes);
        return SFE_FAILURE;
    }
    if (N) {
        if (tree_root_moments(c, pm4, 0, N, theta, false, s)) return SFE_FAILURE;   // same gate on every rank
        const uint32_t s0 = (uint32_t)(((uint64_t)N * pr.rank) / pr.world);
        const uint32_t ns = (uint32_t)(((uint64_t)N * (pr.rank + 1u)) / pr.world) - s0;
        if (ns) {
            WalkRoute rt;
            for (uint32_t r = 0; r < SFE_PEER_MAX; r++) rt.base[r] = (char*)pr.base[r];
            rt.acc_off = acc_off; rt.pot_off = pot_off; rt.world = pr.world; rt.N = N;
            const uint32_t grid = (ns + WALK_THREADS - 1) / WALK_THREADS;
            TreeDev dv = tree_dev(c);
            if (launch_walk<true>(c, flavour, grid, s, dv, pm4, nullptr, s0, ns, G, eps2, theta, nullptr, nullptr, rt)) return SFE_FAILURE;
            if (check_launch(fn)) return SFE_FAILURE;
        }
    }
    return launch_peer_barrier(c, s, fn);
}

}  // namespace sfe

extern "C" int sfe_dev_bh_peer(const void* pm4, uint32_t N, uint64_t acc_off, uint64_t pot_off, float grav_g, float eps2,
                               float theta, uint32_t flavour, void* stream) {
    if (sfe::ensure_init("sfe_dev_bh_peer")) return SFE_FAILURE;
    return sfe::launch_bh_routed(sfe::ctx(), (const float4*)pm4, N, acc_off, pot_off, grav_g, eps2, theta, flavour, (cudaStream_t)stream);
}

extern "C" int sfe_walk_select(int mode) {
    if (mode < -1 || mode > 2) { sfe::set_error("sfe_walk_select", "mode %d is not -1 .. 2", mode); return SFE_FAILURE; }
    sfe::g_walk_select = mode;
    return SFE_SUCCESS;
}

extern "C" int sfe_selftest_exact_math(uint64_t mismatches[3]) {
    using namespace sfe;
    const char* fn = "sfe_selftest_exact_math";
    if (ensure_init(fn)) return SFE_FAILURE;
    Ctx* c = ctx();
    unsigned long long* d = nullptr;
    SFE_CUDA(fn, cudaMalloc(&d, 3 * sizeof(unsigned long long)));
    SFE_CUDA(fn, cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), c->stream));
    exact_math_selftest_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d);
    if (check_launch(fn)) return SFE_FAILURE;
    unsigned long long h[3] = {0, 0, 0};
    SFE_CUDA(fn, cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(fn, cudaStreamSynchronize(c->stream));
    cudaFree(d);
    for (int k = 0; k < 3; k++) mismatches[k] = h[k];
    return SFE_SUCCESS;
}
