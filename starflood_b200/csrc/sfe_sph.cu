// sfe_sph.cu - SPH density, EOS and pressure force (replaces solve_sph, sph.c:14-283).
//
// Reference semantics kept literally: density is an all-pairs Gaussian sum with NO cutoff,
// the force is all-pairs with a hard r^2 > cutoff2 skip, both include j == i, h is per particle.
//
//   rho_i  = sum_j 0.5 (W(r,h_i) + W(r,h_j)) m_j,            W(r,h) = (pi^-1/2 / h)^3 exp(-r^2/h^2)
//   acc_i += -m_i sum_{r^2<=cut} (c_i + c_j) 0.5 (n_i + n_j) r_ij,  n(h) = -2 exp(-r^2/h^2)/(pi^3/2 h^5),
//            c = P/rho^2 (0 if rho^2 == 0)
//
// Both sums are rearranged per SOURCE record so that the pair loop is  3 sub, 3 fma, 2 mul, 2 ex2, few fma:
//   rho_i  = A_i sum_j m_j e_i + sum_j B_j e_j           A = 0.5 a^3,  B_j = 0.5 a_j^3 m_j
//   g_ij   = (c_i + c_j) (D_i e_i + D_j e_j)             D = -1/(pi^3/2 h^5)
// with e_x = exp2(r^2 * k_x), k = -log2(e)/h^2.  When every particle has the same h (the reference never
// updates rad, initcond.c:43-45) e_i == e_j bit for bit and the second MUFU.EX2 is skipped.
//
// Staging is the same TMA bulk-copy / mbarrier double buffer as the PP kernel; sources are 32-byte
// records {x,y,z,q0 | k,q1,-,-}.
#include "sfe_internal.cuh"
#include <stdlib.h>

namespace sfe {

constexpr int SPH_THREADS = 256;
constexpr int SPH_TILE = 256;                  // sources per stage
constexpr int SPH_TILE_BYTES = SPH_TILE * 32;  // 8 KiB
constexpr int SPH_STAGES = 2;
constexpr int SPH_T = 2;                       // targets per thread
constexpr float SPH_PAD_POS = 1.0e15f;

#define SFE_INV_SQRT_PI 0.5641895835477562869480794515607725858440506293289988568440857217
#define SFE_POW_PI_3_2 5.5683279968317078452848179821188357020136243902832439107536758188
#define SFE_LOG2E 1.4426950408889634073599246810018921374266459541529859341354494069

// per-particle coefficients, computed in double and rounded once
__device__ __forceinline__ float coef_k(float h) { return (float)(-SFE_LOG2E / ((double)h * (double)h)); }
__device__ __forceinline__ float coef_a3half(float h) {
    double a = SFE_INV_SQRT_PI / (double)h;
    return (float)(0.5 * a * a * a);
}
__device__ __forceinline__ float coef_d(float h) {
    double h2 = (double)h * (double)h;
    return (float)(-1.0 / (SFE_POW_PI_3_2 * h2 * h2 * (double)h));
}

// mode 0: density records {x,y,z,m | k, B, 0, 0};  mode 1: force records {x,y,z,c | k, D, 0, 0}
__global__ void sph_src_kernel(const float4* __restrict__ pm4, const float* __restrict__ rad,
                               const float* __restrict__ rho, const float* __restrict__ prs, int mode,
                               float4* __restrict__ rec, uint32_t N, uint32_t npadded,
                               uint32_t* __restrict__ h_bits_minmax) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    // h > 0 floats order like their bit patterns: track min / max to detect "all h equal"
    // (one atomic per warp; whole warps are launched, so the full mask is safe)
    {
        const uint32_t hb = j < N ? __float_as_uint(rad[j]) : 0u;
        const uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, j < N ? hb : 0xFFFFFFFFu);
        const uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, hb);
        if ((threadIdx.x & 31u) == 0u && lo != 0xFFFFFFFFu) { atomicMin(&h_bits_minmax[0], lo); atomicMax(&h_bits_minmax[1], hi); }
    }
    if (j >= npadded) return;
    if (j >= N) {
        rec[2u * j] = make_float4(SPH_PAD_POS, SPH_PAD_POS, SPH_PAD_POS, 0.f);
        rec[2u * j + 1u] = make_float4(-1.f, 0.f, 0.f, 0.f);
        return;
    }
    const float4 p = pm4[j];
    const float h = rad[j];
    if (mode == 0) {
        rec[2u * j] = p;
        rec[2u * j + 1u] = make_float4(coef_k(h), coef_a3half(h) * p.w, 0.f, 0.f);
    } else {
        const float r2 = __fmul_rn(rho[j], rho[j]);
        const float cj = 0.f < r2 ? __fdiv_rn(prs[j], r2) : 0.f;   // sph.c:226
        rec[2u * j] = make_float4(p.x, p.y, p.z, cj);
        rec[2u * j + 1u] = make_float4(coef_k(h), coef_d(h), 0.f, 0.f);
    }
}

__global__ void sph_reset_minmax(uint32_t* mm) { mm[0] = 0xFFFFFFFFu; mm[1] = 0u; }

template <int MODE>
__global__ void __launch_bounds__(SPH_THREADS, 2)
sph_pair_kernel(const float4* __restrict__ rec, const float4* __restrict__ pm4, const float* __restrict__ rad,
                const float* __restrict__ rho, const float* __restrict__ prs,
                const uint32_t* __restrict__ h_bits_minmax,
                uint32_t i0, uint32_t ni, uint32_t tiles_total, uint32_t tiles_per_split, float cutoff2,
                float4* __restrict__ partial, const CellGrid* __restrict__ gate) {
    if (gate != nullptr && gate->use_cells) return;   // the cell-list kernels (sfe_sph_cells.cu) took this snapshot
    __shared__ __align__(128) float4 tile[SPH_STAGES][SPH_TILE * 2];
    __shared__ __align__(8) uint64_t full[SPH_STAGES];
    const uint32_t tid = threadIdx.x;
    const uint32_t t0 = blockIdx.y * tiles_per_split;
    const uint32_t t1 = min(t0 + tiles_per_split, tiles_total);
    const uint32_t ntiles = t1 - t0;
    if (tid == 0) {
        for (int s = 0; s < SPH_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t s = 0; s < SPH_STAGES && s < ntiles; s++) {
            mbar_expect_tx(&full[s], SPH_TILE_BYTES);
            bulk_g2s(&tile[s][0], rec + (size_t)(t0 + s) * SPH_TILE * 2, SPH_TILE_BYTES, &full[s]);
        }
    }
    const bool uniform_h = h_bits_minmax[0] == h_bits_minmax[1];

    float xi[SPH_T], yi[SPH_T], zi[SPH_T], ki[SPH_T], qi[SPH_T], ci[SPH_T];
    Kahan tot[SPH_T][3];
    const uint32_t ibase = blockIdx.x * (SPH_THREADS * SPH_T) + tid;
#pragma unroll
    for (int k = 0; k < SPH_T; k++) {
        const uint32_t il = ibase + k * SPH_THREADS;
        const bool ok = il < ni;
        const float4 me = ok ? pm4[i0 + il] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float h = ok ? rad[i0 + il] : 1.f;
        xi[k] = me.x; yi[k] = me.y; zi[k] = me.z;
        ki[k] = coef_k(h);
        if (MODE == 0) { qi[k] = coef_a3half(h); ci[k] = 0.f; }
        else {
            qi[k] = coef_d(h);
            const float r = ok ? rho[i0 + il] : 0.f;
            const float r2 = __fmul_rn(r, r);
            ci[k] = (ok && 0.f < r2) ? __fdiv_rn(prs[i0 + il], r2) : 0.f;   // sph.c:194
        }
#pragma unroll
        for (int q = 0; q < 3; q++) tot[k][q].s = tot[k][q].c = 0.f;
    }

    for (uint32_t it = 0; it < ntiles; it++) {
        const uint32_t s = it % SPH_STAGES;
        mbar_wait(&full[s], (it / SPH_STAGES) & 1u);
        float a0[SPH_T], a1[SPH_T], a2[SPH_T];
#pragma unroll
        for (int k = 0; k < SPH_T; k++) a0[k] = a1[k] = a2[k] = 0.f;
        const float4* sm = &tile[s][0];
        if (uniform_h) {
#pragma unroll 4
            for (int j = 0; j < SPH_TILE; j++) {
                const float4 A = sm[2 * j], B = sm[2 * j + 1];
#pragma unroll
                for (int k = 0; k < SPH_T; k++) {
                    const float dx = xi[k] - A.x, dy = yi[k] - A.y, dz = zi[k] - A.z;
                    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const float e = ex2_approx(r2 * ki[k]);
                    if (MODE == 0) {
                        a0[k] = fmaf(A.w, e, a0[k]);          // sum m_j e_i
                        a1[k] = fmaf(B.y, e, a1[k]);          // sum B_j e_j   (e_j == e_i)
                    } else {
                        float g = (ci[k] + A.w) * ((qi[k] + B.y) * e);
                        g = r2 > cutoff2 ? 0.f : g;            // sph.c:231
                        a0[k] = fmaf(g, dx, a0[k]); a1[k] = fmaf(g, dy, a1[k]); a2[k] = fmaf(g, dz, a2[k]);
                    }
                }
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < SPH_TILE; j++) {
                const float4 A = sm[2 * j], B = sm[2 * j + 1];
#pragma unroll
                for (int k = 0; k < SPH_T; k++) {
                    const float dx = xi[k] - A.x, dy = yi[k] - A.y, dz = zi[k] - A.z;
                    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const float ei = ex2_approx(r2 * ki[k]);
                    const float ej = ex2_approx(r2 * B.x);
                    if (MODE == 0) {
                        a0[k] = fmaf(A.w, ei, a0[k]);
                        a1[k] = fmaf(B.y, ej, a1[k]);
                    } else {
                        float g = (ci[k] + A.w) * fmaf(qi[k], ei, B.y * ej);
                        g = r2 > cutoff2 ? 0.f : g;
                        a0[k] = fmaf(g, dx, a0[k]); a1[k] = fmaf(g, dy, a1[k]); a2[k] = fmaf(g, dz, a2[k]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < SPH_T; k++) {
            if (MODE == 0) { tot[k][0].add(__fmul_rn(qi[k], a0[k])); tot[k][1].add(a1[k]); }
            else { tot[k][0].add(a0[k]); tot[k][1].add(a1[k]); tot[k][2].add(a2[k]); }
        }
        __syncthreads();
        if (tid == 0 && it + SPH_STAGES < ntiles) {
            mbar_expect_tx(&full[s], SPH_TILE_BYTES);
            bulk_g2s(&tile[s][0], rec + (size_t)(t0 + it + SPH_STAGES) * SPH_TILE * 2, SPH_TILE_BYTES, &full[s]);
        }
    }
#pragma unroll
    for (int k = 0; k < SPH_T; k++) {
        const uint32_t il = ibase + k * SPH_THREADS;
        if (il < ni)
            partial[(size_t)blockIdx.y * ni + il] =
                make_float4(tot[k][0].s - tot[k][0].c, tot[k][1].s - tot[k][1].c, tot[k][2].s - tot[k][2].c, 0.f);
    }
}

__global__ void sph_density_finish(const float4* __restrict__ partial, uint32_t i0, uint32_t ni, uint32_t splits,
                                   float* __restrict__ rho, const CellGrid* __restrict__ gate) {
    if (gate != nullptr && gate->use_cells) return;
    uint32_t il = blockIdx.x * blockDim.x + threadIdx.x;
    if (il >= ni) return;
    double s = 0;
    for (uint32_t k = 0; k < splits; k++) { float4 v = partial[(size_t)k * ni + il]; s += (double)v.x + (double)v.y; }
    rho[i0 + il] = (float)s;
}

__global__ void sph_force_finish(const float4* __restrict__ partial, const float4* __restrict__ pm4, uint32_t i0,
                                 uint32_t ni, uint32_t splits, float* __restrict__ acc, const CellGrid* __restrict__ gate) {
    if (gate != nullptr && gate->use_cells) return;
    uint32_t il = blockIdx.x * blockDim.x + threadIdx.x;
    if (il >= ni) return;
    double sx = 0, sy = 0, sz = 0;
    for (uint32_t k = 0; k < splits; k++) { float4 v = partial[(size_t)k * ni + il]; sx += v.x; sy += v.y; sz += v.z; }
    const uint32_t i = i0 + il;
    const float nm = -pm4[i].w;                                  // acc += -m_i * S   (sph.c:274-276)
    acc[3u * i + 0u] = __fadd_rn(acc[3u * i + 0u], __fmul_rn(nm, (float)sx));
    acc[3u * i + 1u] = __fadd_rn(acc[3u * i + 1u], __fmul_rn(nm, (float)sy));
    acc[3u * i + 2u] = __fadd_rn(acc[3u * i + 2u], __fmul_rn(nm, (float)sz));
}

// P = (eos_k rho)^(1/eos_n + 1)  (sph.c:147-159).  The exponent is formed in double and rounded to fp32 like
// the reference; for the reference's constants it is exactly 2 and the compiled reference evaluates x*x
// (gcc folds powf(x, 2.0f)), which is what happens here too.
__global__ void sph_eos_kernel(const float* __restrict__ rho, float* __restrict__ prs, uint32_t n, float eos_k,
                               float exponent) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __fmul_rn(eos_k, rho[i]);
    prs[i] = exponent == 2.0f ? __fmul_rn(x, x) : powf(x, exponent);
}

static void split_plan(const Ctx* c, uint32_t iblocks, uint32_t tiles, uint32_t* splits, uint32_t* tps) {
    const uint32_t wave = 2u * (uint32_t)c->sm_count;
    uint32_t best = 1; double best_eff = 0.0;
    for (uint32_t sp = 1; sp <= 64 && sp <= tiles; sp++) {
        uint32_t t = (tiles + sp - 1) / sp, real_sp = (tiles + t - 1) / t;
        uint64_t ctas = (uint64_t)iblocks * real_sp, waves = (ctas + wave - 1) / wave;
        double eff = (double)iblocks * tiles / ((double)waves * wave * t);
        if (t < 8) eff *= 0.9;
        if (eff > best_eff + 1e-9) { best_eff = eff; best = real_sp; }
    }
    *tps = (tiles + best - 1) / best;
    *splits = (tiles + *tps - 1) / *tps;
}

static int sph_run(Ctx* c, int mode, const float4* pm4, const float* rad, const float* rho, const float* prs,
                   uint32_t N, uint32_t i0, uint32_t ni, float cutoff2, float* out, cudaStream_t s, const char* fn,
                   bool reuse_cells, const SphRoute* route) {
    if (N == 0 || (ni == 0 && !route)) return SFE_SUCCESS;
    if ((uint64_t)i0 + ni > N) { set_error(fn, "target slice [%u,%u) exceeds N=%u", i0, i0 + ni, N); return SFE_FAILURE; }
    // neighbour search first: grid descriptor + sort (skipped when the force pass follows the density pass of
    // the same snapshot inside sfe_sph), then the cell-list pass.  Whether it applies is decided on the device;
    // the all-pairs kernels below are enqueued behind the same flag and retire at once when it does.
    static const bool all_pairs_only = getenv("SFE_SPH_ALL_PAIRS") != nullptr;   // test / measurement hook
    if (!reuse_cells && cells_prepare(c, pm4, rad, N, all_pairs_only, s, fn)) return SFE_FAILURE;
    if (cells_run(c, mode, pm4, rad, rho, prs, N, i0, ni, cutoff2, out, s, fn, route)) return SFE_FAILURE;
    if (ni == 0) return SFE_SUCCESS;   // (a routed rank with an empty index slice still took part in the cell pass)
    const CellGrid* gate = c->sph_cells.maybe ? c->sph_cells.desc.as<CellGrid>() : nullptr;
    const uint32_t tiles = (N + SPH_TILE - 1) / SPH_TILE, npad = tiles * SPH_TILE;
    if (c->sph_src.reserve((size_t)npad * 32 + 64, fn)) return SFE_FAILURE;
    // the last 8 bytes of sph_src hold the (min, max) bit patterns of h, refreshed by every pass
    uint32_t* mm = reinterpret_cast<uint32_t*>(c->sph_src.as<char>() + (size_t)npad * 32);
    sph_reset_minmax<<<1, 1, 0, s>>>(mm);
    if (check_launch(fn)) return SFE_FAILURE;
    sph_src_kernel<<<(npad + 255) / 256, 256, 0, s>>>(pm4, rad, rho, prs, mode, c->sph_src.as<float4>(), N, npad, mm);
    if (check_launch(fn)) return SFE_FAILURE;
    const uint32_t iblocks = (ni + SPH_THREADS * SPH_T - 1) / (SPH_THREADS * SPH_T);
    uint32_t splits, tps;
    split_plan(c, iblocks, tiles, &splits, &tps);
    if (c->sph_partial.reserve((size_t)splits * ni * sizeof(float4), fn)) return SFE_FAILURE;
    dim3 grid(iblocks, splits);
    if (mode == 0) {
        sph_pair_kernel<0><<<grid, SPH_THREADS, 0, s>>>(c->sph_src.as<float4>(), pm4, rad, rho, prs, mm, i0, ni, tiles, tps,
                                                        cutoff2, c->sph_partial.as<float4>(), gate);
        if (check_launch(fn)) return SFE_FAILURE;
        sph_density_finish<<<(ni + 255) / 256, 256, 0, s>>>(c->sph_partial.as<float4>(), i0, ni, splits, out, gate);
    } else {
        sph_pair_kernel<1><<<grid, SPH_THREADS, 0, s>>>(c->sph_src.as<float4>(), pm4, rad, rho, prs, mm, i0, ni, tiles, tps,
                                                        cutoff2, c->sph_partial.as<float4>(), gate);
        if (check_launch(fn)) return SFE_FAILURE;
        sph_force_finish<<<(ni + 255) / 256, 256, 0, s>>>(c->sph_partial.as<float4>(), pm4, i0, ni, splits, out, gate);
    }
    return check_launch(fn);
}

int launch_sph_density(Ctx* c, const float4* pm4, const float* rad, uint32_t N, uint32_t i0, uint32_t ni,
                       float* rho, cudaStream_t s, const SphRoute* route) {
    return sph_run(c, 0, pm4, rad, nullptr, nullptr, N, i0, ni, 0.f, rho, s, "sfe_dev_sph_density", false, route);
}

int launch_sph_eos(const float* rho, float* prs, uint32_t n, float eos_k, float eos_n, cudaStream_t s) {
    if (n == 0) return SFE_SUCCESS;
    const float exponent = (float)((1.0 / (double)eos_n) + 1.0);
    sph_eos_kernel<<<(n + 255) / 256, 256, 0, s>>>(rho, prs, n, eos_k, exponent);
    return check_launch("sfe_dev_sph_eos");
}

int launch_sph_force(Ctx* c, const float4* pm4, const float* rad, const float* rho, const float* prs,
                     uint32_t N, uint32_t i0, uint32_t ni, float cutoff2, float* acc, cudaStream_t s, bool reuse_cells,
                     const SphRoute* route) {
    return sph_run(c, 1, pm4, rad, rho, prs, N, i0, ni, cutoff2, acc, s, "sfe_dev_sph_force", reuse_cells, route);
}

}  // namespace sfe
