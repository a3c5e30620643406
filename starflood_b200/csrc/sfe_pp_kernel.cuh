// sfe_pp_kernel.cuh - the all-pairs gravity kernel template (shared by sfe_pp.cu and tools/pp_tune.cu).
//
// One thread owns T targets; the CTA sweeps a range of j-tiles staged into shared memory by the TMA unit
// (1-D cp.async.bulk completing on an mbarrier, SASS UBLKCP), double buffered.  Sources are PAIR records
// {x0,x1,y0,y1 | z0,z1,m0,m1} so the whole inner loop runs on packed fp32x2 instructions (FADD2 / FFMA2 /
// FMUL2): per interaction 13 FMA-pipe lane-ops + 1 MUFU.RSQ for the 20 flop of the classic count.
//
// The FMA pipe is the bound (20 / 26 = 76.9 % of 2*128*SMs*clk at 100 % pipe utilisation), so everything
// else is arranged to keep it fed:
//   * SMEM_TOT keeps the Kahan-compensated running totals (touched once per tile) in shared memory rather
//     than in registers, which buys the occupancy that hides the MUFU / LDS / tile-boundary stalls;
//   * per-tile partial sums (TILE/2 terms per packed lane) are folded into the compensated totals, and the
//     j-splits are combined in double in a fixed order by pp_finish_kernel: deterministic and within ~1e-6
//     rel-RMS of the reference's per-term Kahan sum (gravity.c:771-799).
#pragma once
#include "sfe_internal.cuh"

namespace sfe {

constexpr int PP_STAGES = 2;
constexpr float PP_PAD_POS = 1.0e15f;        // padding sources: zero mass, far away

template <int T, int THREADS, int MINBLOCKS, int TILE, bool SMEM_TOT, int UNROLL>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
pp_kernel(const float4* __restrict__ pairs,   // pair records, whole tiles
          const float4* __restrict__ pm4,     // targets are read from here
          uint32_t N, uint32_t i0, uint32_t ni, uint32_t tiles_total, uint32_t tiles_per_split, float eps2,
          float4* __restrict__ partial /* [gridDim.y][ni] */) {
    constexpr int TILE_BYTES = TILE * 16;
    __shared__ __align__(128) float4 tile[PP_STAGES][TILE];  // TILE/2 pairs x 2 float4
    __shared__ __align__(8) uint64_t full[PP_STAGES];
    __shared__ float stot[SMEM_TOT ? T * 8 : 1][SMEM_TOT ? THREADS : 1];

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
            mbar_expect_tx(&full[s], TILE_BYTES);
            bulk_g2s(&tile[s][0], pairs + (size_t)(t0 + s) * TILE, TILE_BYTES, &full[s]);
        }
    }

    // targets: negated so that d' = r_j - r_i is one FADD2 (the sign is undone at the end)
    float2 nx[T], ny[T], nz[T];
    Kahan tot[SMEM_TOT ? 1 : T][4];
    const uint32_t ibase = blockIdx.x * (THREADS * T) + tid;
#pragma unroll
    for (int k = 0; k < T; k++) {
        uint32_t il = ibase + k * THREADS;
        float4 me = (il < ni) ? pm4[i0 + il] : make_float4(0.f, 0.f, 0.f, 0.f);
        nx[k] = make_float2(-me.x, -me.x);
        ny[k] = make_float2(-me.y, -me.y);
        nz[k] = make_float2(-me.z, -me.z);
        if (SMEM_TOT) {
#pragma unroll
            for (int q = 0; q < 8; q++) stot[k * 8 + q][tid] = 0.f;
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) tot[k][q].s = tot[k][q].c = 0.f;
        }
    }
    const float2 e2 = make_float2(eps2, eps2);

    for (uint32_t it = 0; it < ntiles; it++) {
        const uint32_t s = it % PP_STAGES;
        mbar_wait(&full[s], (it / PP_STAGES) & 1u);
        float2 ax[T], ay[T], az[T], ap[T];
#pragma unroll
        for (int k = 0; k < T; k++) ax[k] = ay[k] = az[k] = ap[k] = make_float2(0.f, 0.f);
        const float4* sm = &tile[s][0];
#pragma unroll UNROLL
        for (int p = 0; p < TILE / 2; p++) {
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
            const float v[4] = {ax[k].x + ax[k].y, ay[k].x + ay[k].y, az[k].x + az[k].y, ap[k].x + ap[k].y};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (SMEM_TOT) {
                    Kahan a{stot[k * 8 + 2 * q][tid], stot[k * 8 + 2 * q + 1][tid]};
                    a.add(v[q]);
                    stot[k * 8 + 2 * q][tid] = a.s; stot[k * 8 + 2 * q + 1][tid] = a.c;
                } else {
                    tot[k][q].add(v[q]);
                }
            }
        }
        __syncthreads();  // everyone is done reading stage s
        if (tid == 0 && it + PP_STAGES < ntiles) {
            mbar_expect_tx(&full[s], TILE_BYTES);
            bulk_g2s(&tile[s][0], pairs + (size_t)(t0 + it + PP_STAGES) * TILE, TILE_BYTES, &full[s]);
        }
    }

#pragma unroll
    for (int k = 0; k < T; k++) {
        uint32_t il = ibase + k * THREADS;
        if (il < ni) {
            float r[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
                r[q] = SMEM_TOT ? stot[k * 8 + 2 * q][tid] - stot[k * 8 + 2 * q + 1][tid] : tot[SMEM_TOT ? 0 : k][q].s - tot[SMEM_TOT ? 0 : k][q].c;
            // d' = -(r_i - r_j): flip the sign of the force sums here; the compensation terms are folded in
            partial[(size_t)blockIdx.y * ni + il] = make_float4(-r[0], -r[1], -r[2], r[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Target-pair variant: the two packed lanes are two TARGETS, the source is a broadcast scalar record
// {x, y, z, m} (16 B, one LDS.128 per source).  Same 13 FMA-pipe lane-ops + 1 MUFU per interaction, but half
// the shared-memory -> register traffic per interaction of the source-pair kernel above (the LDS write-backs
// compete with the FMA pipe for register-file ports), and no source re-packing pass: tiles are bulk-copied
// straight from the {x,y,z,m} array the all-gather produces.  Thread t of a CTA owns targets
// i = blockIdx.x * THREADS * 2T + k * 2 * THREADS + {t, t + THREADS}.
template <int T, int THREADS, int TILE>
constexpr size_t pp_tp_smem_bytes() { return (size_t)PP_STAGES * TILE * 16 + (size_t)T * 16 * THREADS * 4 + PP_STAGES * 8; }

// ACC: 0 = accumulate with FFMA2 (three fresh 64-bit operands), 1 = two scalar FFMA per accumulate (three 32-bit
// operands each run at full rate, the packed form is limited by register-file operand bandwidth).
// UM: every source has the same mass (checked by the launcher): the mass factor moves out of the sums.
template <int T, int THREADS, int MINBLOCKS, int TILE, int UNROLL, int ACC = 0, bool UM = false>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
pp_kernel_tp(const float4* __restrict__ src,     // sources {x,y,z,m}, padded to whole tiles
             const float4* __restrict__ pm4,     // targets are read from here
             uint32_t N, uint32_t i0, uint32_t ni, uint32_t tiles_total, uint32_t tiles_per_split, float eps2,
             float4* __restrict__ partial /* [gridDim.y][ni] */,
             const uint32_t* __restrict__ mixed_masses = nullptr /* device flag written by pp_mass_scan_kernel */) {
    // the launcher enqueues both specialisations; the one that does not match the flag retires immediately
    if (mixed_masses != nullptr && (*mixed_masses != 0u) == UM) return;
    constexpr int TILE_BYTES = TILE * 16;
    extern __shared__ __align__(128) unsigned char pp_smem[];   // pp_tp_smem_bytes<T, THREADS, TILE>()
    float4 (*tile)[TILE] = reinterpret_cast<float4 (*)[TILE]>(pp_smem);                                 // [PP_STAGES][TILE]
    float (*stot)[THREADS] = reinterpret_cast<float (*)[THREADS]>(pp_smem + PP_STAGES * TILE_BYTES);   // [T * 16][THREADS]
    uint64_t* full = reinterpret_cast<uint64_t*>(pp_smem + PP_STAGES * TILE_BYTES + T * 16 * THREADS * 4);
    // stot: Kahan (sum, comp) x {x,y,z,pot} x 2 targets per pair

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
            mbar_expect_tx(&full[s], TILE_BYTES);
            bulk_g2s(&tile[s][0], src + (size_t)(t0 + s) * TILE, TILE_BYTES, &full[s]);
        }
    }

    float2 nx[T], ny[T], nz[T];   // negated target coordinates, {target a, target b}
    const uint32_t ibase = blockIdx.x * (THREADS * 2 * T) + tid;
#pragma unroll
    for (int k = 0; k < T; k++) {
        const uint32_t ia = ibase + k * 2 * THREADS, ib = ia + THREADS;
        const float4 a = (ia < ni) ? pm4[i0 + ia] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 b = (ib < ni) ? pm4[i0 + ib] : make_float4(0.f, 0.f, 0.f, 0.f);
        nx[k] = make_float2(-a.x, -b.x);
        ny[k] = make_float2(-a.y, -b.y);
        nz[k] = make_float2(-a.z, -b.z);
#pragma unroll
        for (int q = 0; q < 16; q++) stot[k * 16 + q][tid] = 0.f;
    }
    const float2 e2 = make_float2(eps2, eps2);

    for (uint32_t it = 0; it < ntiles; it++) {
        const uint32_t s = it % PP_STAGES;
        mbar_wait(&full[s], (it / PP_STAGES) & 1u);
        float2 ax[T], ay[T], az[T], ap[T];
#pragma unroll
        for (int k = 0; k < T; k++) ax[k] = ay[k] = az[k] = ap[k] = make_float2(0.f, 0.f);
        const float4* sm = &tile[s][0];
#pragma unroll UNROLL
        for (int j = 0; j < TILE; j++) {
            const float4 S = sm[j];
            const float2 xj = make_float2(S.x, S.x), yj = make_float2(S.y, S.y);
            const float2 zj = make_float2(S.z, S.z), mj = make_float2(S.w, S.w);
#pragma unroll
            for (int k = 0; k < T; k++) {
                const float2 dx = __fadd2_rn(nx[k], xj);
                const float2 dy = __fadd2_rn(ny[k], yj);
                const float2 dz = __fadd2_rn(nz[k], zj);
                float2 r2 = __ffma2_rn(dx, dx, e2);
                r2 = __ffma2_rn(dy, dy, r2);
                r2 = __ffma2_rn(dz, dz, r2);
                float2 inv;
                inv.x = rsqrt_approx(r2.x);
                inv.y = rsqrt_approx(r2.y);
                const float2 u = UM ? inv : __fmul2_rn(inv, mj);     // m_j / r          (pot_j, gravity.c:761)
                const float2 i2 = __fmul2_rn(inv, inv);   // 1 / r^2
                const float2 w = __fmul2_rn(u, i2);       // m_j / r^3
                if (ACC == 0) {
                    ax[k] = __ffma2_rn(w, dx, ax[k]);
                    ay[k] = __ffma2_rn(w, dy, ay[k]);
                    az[k] = __ffma2_rn(w, dz, az[k]);
                } else {
                    ax[k].x = __fmaf_rn(w.x, dx.x, ax[k].x); ax[k].y = __fmaf_rn(w.y, dx.y, ax[k].y);
                    ay[k].x = __fmaf_rn(w.x, dy.x, ay[k].x); ay[k].y = __fmaf_rn(w.y, dy.y, ay[k].y);
                    az[k].x = __fmaf_rn(w.x, dz.x, az[k].x); az[k].y = __fmaf_rn(w.y, dz.y, az[k].y);
                }
                ap[k] = __fadd2_rn(ap[k], u);
            }
        }
        // fold this tile's partial sums (TILE terms per lane) into the compensated running totals
#pragma unroll
        for (int k = 0; k < T; k++) {
            const float v[8] = {ax[k].x, ay[k].x, az[k].x, ap[k].x, ax[k].y, ay[k].y, az[k].y, ap[k].y};
#pragma unroll
            for (int q = 0; q < 8; q++) {
                Kahan a{stot[k * 16 + 2 * q][tid], stot[k * 16 + 2 * q + 1][tid]};
                a.add(v[q]);
                stot[k * 16 + 2 * q][tid] = a.s; stot[k * 16 + 2 * q + 1][tid] = a.c;
            }
        }
        __syncthreads();  // everyone is done reading stage s
        if (tid == 0 && it + PP_STAGES < ntiles) {
            mbar_expect_tx(&full[s], TILE_BYTES);
            bulk_g2s(&tile[s][0], src + (size_t)(t0 + it + PP_STAGES) * TILE, TILE_BYTES, &full[s]);
        }
    }

#pragma unroll
    for (int k = 0; k < T; k++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t il = ibase + k * 2 * THREADS + h * THREADS;
            if (il < ni) {
                float r[4];
#pragma unroll
                for (int q = 0; q < 4; q++) r[q] = stot[k * 16 + 8 * h + 2 * q][tid] - stot[k * 16 + 8 * h + 2 * q + 1][tid];
                // d' = -(r_i - r_j): flip the sign of the force sums here
                partial[(size_t)blockIdx.y * ni + il] = make_float4(-r[0], -r[1], -r[2], r[3]);
            }
        }
    }
}

}  // namespace sfe
