// sfe_pp.cu - all-pairs softened gravity (replaces solve_gravity_part_part, gravity.c:667-819).
//
// Bound: FP32 FMA pipe (20 flop per (i,j) interaction, N^2 interactions, j == i included).
//
// Design (sm_100a):
//   * sources are repacked once per solve into PAIRS  {x0,x1,y0,y1 | z0,z1,m0,m1}  (32 B per two
//     sources) so that the inner loop can run entirely on Blackwell's packed fp32x2 instructions
//     (FADD2 / FFMA2 / FMUL2, sm_100+): one thread evaluates one target against TWO sources per
//     instruction, which halves the issue slots per interaction and leaves room for the MUFU.RSQ and
//     the shared-memory loads next to an FMA pipe that stays busy;
//   * j-tiles of 512 sources (8 KiB) are staged into shared memory by the TMA unit with 1-D bulk
//     copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier, double buffered, issued by one
//     elected thread; every lane of a warp reads the same pair -> LDS.128 broadcasts;
//   * i-blocking: each thread owns T targets in registers, so one pair of LDS.128 feeds 2*T interactions;
//   * per interaction: 3 sub, 3 fma (r^2 + eps^2), 1 rsqrt, 3 mul, 3 fma (acc), 1 add (pot) = 13 FMA-pipe
//     ops + 1 MUFU for the 20 flop of the classic count;
//   * accuracy without per-term Kahan: partial sums over one tile (256 terms per packed lane) are folded
//     into Kahan-compensated running totals, and j-splits are combined in double in a fixed order, so
//     the result is deterministic and within ~1e-6 rel-RMS of the reference's compensated fp32 sum;
//   * grid: (target blocks) x (j splits) chosen so that the CTA count fills 148 SMs x 2 CTAs evenly even
//     at N = 65536; partials go to a [split][i] float4 buffer and a second kernel applies -G m_i.
#include "sfe_internal.cuh"

namespace sfe {

constexpr int PP_THREADS = 256;
constexpr int PP_TILE = 512;                 // sources per shared-memory stage
constexpr int PP_TILE_BYTES = PP_TILE * 16;  // 8 KiB
constexpr int PP_STAGES = 2;
constexpr float PP_PAD_POS = 1.0e15f;        // padding sources: zero mass, far away

// ---------------------------------------------------------------- packing kernels
__global__ void pack_pm4_kernel(const float* __restrict__ pos, const float* __restrict__ mas,
                                float4* __restrict__ pm4, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pm4[i] = make_float4(pos[3u * i], pos[3u * i + 1u], pos[3u * i + 2u], mas[i]);
}

int launch_pack_pm4(const float* pos, const float* mas, float4* pm4, uint32_t n, cudaStream_t s) {
    if (n == 0) return SFE_SUCCESS;
    pack_pm4_kernel<<<(n + 255) / 256, 256, 0, s>>>(pos, mas, pm4, n);
    return check_launch("sfe_dev_pack_pm4");
}

// pm4 {x,y,z,m} -> pair records {x0,x1,y0,y1}{z0,z1,m0,m1}; pads to a whole number of tiles.
__global__ void pp_pair_kernel(const float4* __restrict__ pm4, float4* __restrict__ pairs, uint32_t N,
                               uint32_t npairs_padded) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs_padded) return;
    const float4 pad = make_float4(PP_PAD_POS, PP_PAD_POS, PP_PAD_POS, 0.0f);
    float4 a = (2u * p < N) ? pm4[2u * p] : pad;
    float4 b = (2u * p + 1u < N) ? pm4[2u * p + 1u] : pad;
    pairs[2u * p] = make_float4(a.x, b.x, a.y, b.y);
    pairs[2u * p + 1u] = make_float4(a.z, b.z, a.w, b.w);
}

// ---------------------------------------------------------------- main kernel

template <int T>
__global__ void __launch_bounds__(PP_THREADS, 2)
pp_kernel(const float4* __restrict__ pairs,   // pair records, whole tiles
          const float4* __restrict__ pm4,     // targets are read from here
          uint32_t N, uint32_t i0, uint32_t ni, uint32_t tiles_total, uint32_t tiles_per_split, float eps2,
          float4* __restrict__ partial /* [gridDim.y][ni] */) {
    __shared__ __align__(128) float4 tile[PP_STAGES][PP_TILE];  // PP_TILE/2 pairs x 2 float4
    __shared__ __align__(8) uint64_t full[PP_STAGES];

    const uint32_t tid = threadIdx.x;
    const uint32_t t0 = blockIdx.y * tiles_per_split;
    const uint32_t t1 = min(t0 + tiles_per_split, tiles_total);
    const uint32_t ntiles = t1 - t0;

    if (tid == 0) {
        for (int s = 0; s < PP_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t s = 0; s < PP_STAGES && s < ntiles; s++) {
            mbar_expect_tx(&full[s], PP_TILE_BYTES);
            bulk_g2s(&tile[s][0], pairs + (size_t)(t0 + s) * PP_TILE, PP_TILE_BYTES, &full[s]);
        }
    }

    // targets: negated and duplicated so that d' = r_j - r_i is one FADD2 (the sign is undone at the end)
    float2 nx[T], ny[T], nz[T];
    Kahan tot[T][4];
    const uint32_t ibase = blockIdx.x * (PP_THREADS * T) + tid;
#pragma unroll
    for (int k = 0; k < T; k++) {
        uint32_t il = ibase + k * PP_THREADS;
        float4 me = (il < ni) ? pm4[i0 + il] : make_float4(0.f, 0.f, 0.f, 0.f);
        nx[k] = make_float2(-me.x, -me.x);
        ny[k] = make_float2(-me.y, -me.y);
        nz[k] = make_float2(-me.z, -me.z);
#pragma unroll
        for (int q = 0; q < 4; q++) tot[k][q].s = tot[k][q].c = 0.f;
    }
    const float2 e2 = make_float2(eps2, eps2);

    for (uint32_t it = 0; it < ntiles; it++) {
        const uint32_t s = it % PP_STAGES;
        mbar_wait(&full[s], (it / PP_STAGES) & 1u);
        float2 ax[T], ay[T], az[T], ap[T];
#pragma unroll
        for (int k = 0; k < T; k++) ax[k] = ay[k] = az[k] = ap[k] = make_float2(0.f, 0.f);
        const float4* sm = &tile[s][0];
#pragma unroll 4
        for (int p = 0; p < PP_TILE / 2; p++) {
            const float4 A = sm[2 * p];      // x0 x1 y0 y1
            const float4 B = sm[2 * p + 1];  // z0 z1 m0 m1
            const float2 xj = make_float2(A.x, A.y), yj = make_float2(A.z, A.w);
            const float2 zj = make_float2(B.x, B.y), mj = make_float2(B.z, B.w);
#pragma unroll
            for (int k = 0; k < T; k++) {
                const float2 dx = __fadd2_rn(xj, nx[k]);
                const float2 dy = __fadd2_rn(yj, ny[k]);
                const float2 dz = __fadd2_rn(zj, nz[k]);
                float2 r2 = __ffma2_rn(dx, dx, e2);
                r2 = __ffma2_rn(dy, dy, r2);
                r2 = __ffma2_rn(dz, dz, r2);
                float2 inv;
                inv.x = rsqrt_approx(r2.x);
                inv.y = rsqrt_approx(r2.y);
                const float2 u = __fmul2_rn(mj, inv);     // m_j / r          (pot_j, gravity.c:761)
                const float2 i2 = __fmul2_rn(inv, inv);   // 1 / r^2
                const float2 w = __fmul2_rn(u, i2);       // m_j / r^3
                ax[k] = __ffma2_rn(w, dx, ax[k]);
                ay[k] = __ffma2_rn(w, dy, ay[k]);
                az[k] = __ffma2_rn(w, dz, az[k]);
                ap[k] = __fadd2_rn(ap[k], u);
            }
        }
        // fold this tile's partial sums into the compensated running totals
#pragma unroll
        for (int k = 0; k < T; k++) {
            tot[k][0].add(ax[k].x + ax[k].y);
            tot[k][1].add(ay[k].x + ay[k].y);
            tot[k][2].add(az[k].x + az[k].y);
            tot[k][3].add(ap[k].x + ap[k].y);
        }
        __syncthreads();  // everyone is done reading stage s
        if (tid == 0 && it + PP_STAGES < ntiles) {
            mbar_expect_tx(&full[s], PP_TILE_BYTES);
            bulk_g2s(&tile[s][0], pairs + (size_t)(t0 + it + PP_STAGES) * PP_TILE, PP_TILE_BYTES, &full[s]);
        }
    }

#pragma unroll
    for (int k = 0; k < T; k++) {
        uint32_t il = ibase + k * PP_THREADS;
        if (il < ni) {
            // d' = -(r_i - r_j): flip the sign of the force sums here; fold in the compensation terms
            partial[(size_t)blockIdx.y * ni + il] =
                make_float4(-(tot[k][0].s - tot[k][0].c), -(tot[k][1].s - tot[k][1].c),
                            -(tot[k][2].s - tot[k][2].c), tot[k][3].s - tot[k][3].c);
        }
    }
}

// combine the j-splits in a fixed order (double accumulation) and apply the reference's prefactors
// acc = ((-G) * m_i) * S, pot = (-G) * Phi   (gravity.c:808-811)
__global__ void pp_finish_kernel(const float4* __restrict__ partial, const float4* __restrict__ pm4,
                                 uint32_t i0, uint32_t ni, uint32_t splits, float G,
                                 float* __restrict__ acc, float* __restrict__ pot) {
    uint32_t il = blockIdx.x * blockDim.x + threadIdx.x;
    if (il >= ni) return;
    double sx = 0, sy = 0, sz = 0, sp = 0;
    for (uint32_t s = 0; s < splits; s++) {
        float4 v = partial[(size_t)s * ni + il];
        sx += v.x; sy += v.y; sz += v.z; sp += v.w;
    }
    const uint32_t i = i0 + il;
    const float gm = __fmul_rn(-G, pm4[i].w);
    acc[3u * i + 0u] = __fmul_rn(gm, (float)sx);
    acc[3u * i + 1u] = __fmul_rn(gm, (float)sy);
    acc[3u * i + 2u] = __fmul_rn(gm, (float)sz);
    pot[i] = __fmul_rn(-G, (float)sp);
}

// ---------------------------------------------------------------- launcher
static int pp_targets_per_thread() {
    static int t = -1;
    if (t < 0) {
        const char* e = getenv("SFE_PP_T");
        t = e ? atoi(e) : 4;
        if (t != 1 && t != 2 && t != 4) t = 4;
    }
    return t;
}

int launch_pp(Ctx* c, const float4* pm4, uint32_t N, uint32_t i0, uint32_t ni, float* acc, float* pot,
              float G, float eps2, cudaStream_t s) {
    const char* fn = "sfe_dev_pp";
    if (ni == 0 || N == 0) return SFE_SUCCESS;
    if ((uint64_t)i0 + ni > N) { set_error(fn, "target slice [%u,%u) exceeds N=%u", i0, i0 + ni, N); return SFE_FAILURE; }
    const uint32_t tiles = (N + PP_TILE - 1) / PP_TILE;
    const uint32_t npairs = tiles * (PP_TILE / 2);
    if (c->pp_src.reserve((size_t)npairs * 32, fn)) return SFE_FAILURE;
    pp_pair_kernel<<<(npairs + 255) / 256, 256, 0, s>>>(pm4, c->pp_src.as<float4>(), N, npairs);
    if (check_launch(fn)) return SFE_FAILURE;

    const int T = pp_targets_per_thread();
    const uint32_t iblocks = (ni + PP_THREADS * T - 1) / (PP_THREADS * T);
    // choose the j-split count so that the grid is close to a whole number of waves of 2 CTAs per SM
    const uint32_t wave = 2u * (uint32_t)c->sm_count;
    uint32_t splits = 1;
    if (iblocks < 8u * wave) {
        uint32_t best = 1; double best_eff = 0.0;
        for (uint32_t sp = 1; sp <= 64 && sp <= tiles; sp++) {
            uint32_t tps = (tiles + sp - 1) / sp;
            uint32_t real_sp = (tiles + tps - 1) / tps;
            uint64_t ctas = (uint64_t)iblocks * real_sp;
            uint64_t waves = (ctas + wave - 1) / wave;
            // efficiency: useful tile-work / (waves * wave * tiles-per-split)
            double eff = (double)iblocks * tiles / ((double)waves * wave * tps);
            if (tps < 8) eff *= 0.9;  // very short splits pay prologue/epilogue
            if (eff > best_eff + 1e-9) { best_eff = eff; best = real_sp; }
        }
        splits = best;
    }
    const uint32_t tps = (tiles + splits - 1) / splits;
    splits = (tiles + tps - 1) / tps;
    if (c->pp_partial.reserve((size_t)splits * ni * sizeof(float4), fn)) return SFE_FAILURE;

    dim3 grid(iblocks, splits);
    switch (T) {
        case 1: pp_kernel<1><<<grid, PP_THREADS, 0, s>>>(c->pp_src.as<float4>(), pm4, N, i0, ni, tiles, tps, eps2, c->pp_partial.as<float4>()); break;
        case 2: pp_kernel<2><<<grid, PP_THREADS, 0, s>>>(c->pp_src.as<float4>(), pm4, N, i0, ni, tiles, tps, eps2, c->pp_partial.as<float4>()); break;
        default: pp_kernel<4><<<grid, PP_THREADS, 0, s>>>(c->pp_src.as<float4>(), pm4, N, i0, ni, tiles, tps, eps2, c->pp_partial.as<float4>()); break;
    }
    if (check_launch(fn)) return SFE_FAILURE;
    pp_finish_kernel<<<(ni + 255) / 256, 256, 0, s>>>(c->pp_partial.as<float4>(), pm4, i0, ni, splits, G, acc, pot);
    return check_launch(fn);
}

}  // namespace sfe
