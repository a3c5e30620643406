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
// stays mostly converged.  The 32-byte decision record {centre, width^2, next, child|kind} is the only thing
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
// Bound: instruction issue (EXACT: 1.6e9 warp instructions per launch at N = 2^20, issue slots 82 % busy, 21 of 32 lanes
// active; FAST: L1 sector look-ups of divergent node loads) - the node table stays L1/L2 resident, DRAM traffic is 0.4 % of
// the bytes of SURVEY 8(d)'s HBM model (32 B x node visits + particle I/O), which bench.py reports beside the binding roof.
#include "sfe_internal.cuh"
#include "sfe_tree.cuh"
#include "sfe_scan.cuh"

namespace sfe {

constexpr int WALK_THREADS = 128;

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
    }
};

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

    KahanRef2 kxy = {{0.f, 0.f}, {0.f, 0.f}}, kzp = {{0.f, 0.f}, {0.f, 0.f}};   // (acc x, acc y) and (acc z, pot)
    float fx = 0.f, fy = 0.f, fz = 0.f, fp = 0.f;

    const WalkRec0* __restrict__ rec0 = t.rec0;
    const float4* __restrict__ rec1 = t.rec1;
    const WalkRec2* __restrict__ rec2 = t.rec2;

    const bool seq_ok = eps2 >= SEQ_LO;
    uint32_t n = M == 0u ? NODE_END : 0u;   // an empty root is never evaluated (gravity.c:324-328)
    while (n != NODE_END) {
        // one 256-bit load = one L1 sector: {a[4], next, child|kind, rin2, -}
        uint32_t r[8];
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "l"(rec0 + n));
        const float a0 = __uint_as_float(r[0]), a1 = __uint_as_float(r[1]), a2 = __uint_as_float(r[2]), a3 = __uint_as_float(r[3]);
        const uint32_t next = r[4], kind = r[5] >> 30;
        const float rin2 = __uint_as_float(r[6]);
        float rx, ry, rz, mj;
        bool maybe_inside;
        if (kind == KIND_INTERNAL) {
            const float2 oxy = __fadd2_rn(make_float2(a0, a1), make_float2(-x, -y));   // packed: same roundings, fewer issue slots
            const float2 osq = __fmul2_rn(oxy, oxy);
            const float oz = __fsub_rn(a2, z);
            const float d2 = __fadd_rn(__fadd_rn(osq.x, osq.y), __fmul_rn(oz, oz));
            if (opens(a3, d2, theta)) { n = r[5] & 0x3FFFFFFFu; continue; }   // gravity.c:448-453 (first non-empty child)
            const float4 cm = __ldg(&rec1[n]);                                   // accepted: monopole of the cell
            rx = cm.x; ry = cm.y; rz = cm.z; mj = cm.w;
            maybe_inside = d2 <= rin2;
        } else {                                                                 // leaf: the record IS the monopole
            rx = a0; ry = a1; rz = a2; mj = a3;
            const float ux = x - rx, uy = y - ry, uz = z - rz;
            maybe_inside = kind != KIND_SINGLE && (ux * ux + uy * uy + uz * uz) <= rin2;
        }
        // monopole (gravity.c:463-548)
        if (maybe_inside) {
            const float4 lo = __ldg(reinterpret_cast<const float4*>(&rec2[n]));
            const float4 hi = __ldg(reinterpret_cast<const float4*>(&rec2[n]) + 1);
            // lo = {lo.x, lo.y, lo.z, hi.x}, hi = {hi.y, hi.z, -, -}
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
    if (flavour == SFE_WALK_EXACT)
        walk_kernel<(int)SFE_WALK_EXACT, false><<<grid, WALK_THREADS, 0, s>>>(dv, pm4, slots, 0u, ni, G, eps2, theta, acc, pot, none);
    else
        walk_kernel<(int)SFE_WALK_FAST, false><<<grid, WALK_THREADS, 0, s>>>(dv, pm4, slots, 0u, ni, G, eps2, theta, acc, pot, none);
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
            if (flavour == SFE_WALK_EXACT)
                walk_kernel<(int)SFE_WALK_EXACT, true><<<grid, WALK_THREADS, 0, s>>>(dv, pm4, nullptr, s0, ns, G, eps2, theta, nullptr, nullptr, rt);
            else
                walk_kernel<(int)SFE_WALK_FAST, true><<<grid, WALK_THREADS, 0, s>>>(dv, pm4, nullptr, s0, ns, G, eps2, theta, nullptr, nullptr, rt);
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
