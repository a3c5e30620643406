// sfe_pp.cu - all-pairs softened gravity (replaces solve_gravity_part_part, gravity.c:667-819).
//
// Bound: FP32 FMA pipe (20 flop per (i,j) interaction, N^2 interactions, j == i included).
//
// Design (sm_100a), kernel template in sfe_pp_kernel.cuh:
//   * the inner loop runs entirely on Blackwell's packed fp32x2 instructions (FADD2 / FFMA2 / FMUL2, sm_100+):
//     the two packed lanes are two TARGETS, the source {x,y,z,m} is a broadcast scalar operand, so one
//     instruction evaluates two interactions and the issue slots left over hide the MUFU.RSQ and the LDS;
//   * j-tiles of 512 sources (8 KiB) are staged into shared memory by the TMA unit with 1-D bulk copies
//     (cp.async.bulk, SASS UBLKCP) straight from the {x,y,z,m} array (the unit the all-gather moves),
//     completing on an mbarrier, double buffered, issued by one elected thread; every lane of a warp reads
//     the same source -> one broadcast LDS.128 per source;
//   * i-blocking: each thread owns PP_T target pairs in registers, so one LDS.128 feeds 2*PP_T interactions;
//   * per interaction: 3 sub, 3 fma (r^2 + eps^2), 1 rsqrt, 3 mul, 3 fma (acc), 1 add (pot) = 13 FMA-pipe
//     ops + 1 MUFU for the 20 flop of the classic count -> ceiling 20/26 = 76.9 % of 2*128*SMs*clk;
//   * accuracy without per-term Kahan: partial sums over one tile (512 terms) are folded into
//     Kahan-compensated running totals kept in shared memory, and j-splits are combined in double in a fixed
//     order, so the result is deterministic and within ~1e-6 rel-RMS of the reference's compensated fp32 sum;
//   * grid: (target blocks) x (j splits) chosen so that the CTA count is a near-whole number of waves of
//     148 SMs x resident CTAs even at N = 65536; partials go to a [split][i] float4 buffer and a second kernel
//     applies -G m_i.
// Variant choice (T, CTA size, occupancy, tile) was measured with tools/pp_tune.cu (profiles/r01_pp_tune*.txt).
#include "sfe_internal.cuh"
#include "sfe_pp_kernel.cuh"
#include <algorithm>

namespace sfe {

constexpr int PP_T = 6;                      // target pairs per thread (12 targets)
constexpr int PP_THREADS = 128;
constexpr int PP_MINB = 2;
constexpr int PP_TILE = 512;                 // sources per shared-memory stage
constexpr int PP_UNROLL = 2;

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

// sources padded to a whole number of tiles (zero mass, far away); only needed when N % PP_TILE != 0
__global__ void pp_pad_kernel(const float4* __restrict__ pm4, float4* __restrict__ src, uint32_t N, uint32_t npad) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < npad) src[j] = j < N ? pm4[j] : make_float4(PP_PAD_POS, PP_PAD_POS, PP_PAD_POS, 0.0f);
}

// Equal-mass detection: flag = 1 as soon as one source's mass differs (bitwise) from source 0's.  Runs on the
// stream ahead of the pair kernels, which read the flag on the device (no host round trip).
__global__ void pp_mass_scan_kernel(const float4* __restrict__ pm4, uint32_t N, uint32_t* __restrict__ flag) {
    const uint32_t m0 = __float_as_uint(pm4[0].w);
    bool differs = false;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x)
        differs |= __float_as_uint(pm4[j].w) != m0;
    if (__syncthreads_or(differs) && threadIdx.x == 0) atomicOr(flag, 1u);
}

// combine the j-splits in a fixed order (double accumulation) and apply the reference's prefactors
// acc = ((-G) * m_i) * S, pot = (-G) * Phi   (gravity.c:808-811)
__global__ void pp_finish_kernel(const float4* __restrict__ partial, const float4* __restrict__ pm4,
                                 uint32_t i0, uint32_t ni, uint32_t splits, float G,
                                 float* __restrict__ acc, float* __restrict__ pot,
                                 const uint32_t* __restrict__ mixed_masses) {
    uint32_t il = blockIdx.x * blockDim.x + threadIdx.x;
    if (il >= ni) return;
    double sx = 0, sy = 0, sz = 0, sp = 0;
    for (uint32_t s = 0; s < splits; s++) {
        float4 v = partial[(size_t)s * ni + il];
        sx += v.x; sy += v.y; sz += v.z; sp += v.w;
    }
    if (*mixed_masses == 0u) {   // equal-mass specialisation: the common m_j was left out of the sums
        const double mj = (double)pm4[0].w;
        sx *= mj; sy *= mj; sz *= mj; sp *= mj;
    }
    const uint32_t i = i0 + il;
    const float gm = __fmul_rn(-G, pm4[i].w);
    acc[3u * i + 0u] = __fmul_rn(gm, (float)sx);
    acc[3u * i + 1u] = __fmul_rn(gm, (float)sy);
    acc[3u * i + 2u] = __fmul_rn(gm, (float)sz);
    pot[i] = __fmul_rn(-G, (float)sp);
}

// ---------------------------------------------------------------- launcher
int launch_pp(Ctx* c, const float4* pm4, uint32_t N, uint32_t i0, uint32_t ni, float* acc, float* pot,
              float G, float eps2, cudaStream_t s) {
    const char* fn = "sfe_dev_pp";
    if (ni == 0 || N == 0) return SFE_SUCCESS;
    if ((uint64_t)i0 + ni > N) { set_error(fn, "target slice [%u,%u) exceeds N=%u", i0, i0 + ni, N); return SFE_FAILURE; }
    auto kernel = pp_kernel_tp<PP_T, PP_THREADS, PP_MINB, PP_TILE, PP_UNROLL, 0, false>;   // general masses
    auto kernel_um = pp_kernel_tp<PP_T, PP_THREADS, PP_MINB, PP_TILE, PP_UNROLL, 0, true>;  // all m_j equal
    constexpr size_t smem = pp_tp_smem_bytes<PP_T, PP_THREADS, PP_TILE>();
    static int occ = 0;
    if (occ == 0) {
        SFE_CUDA(fn, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SFE_CUDA(fn, cudaFuncSetAttribute(kernel_um, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SFE_CUDA(fn, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, PP_THREADS, smem));
        if (occ < 1) occ = 1;
    }
    const uint32_t tiles = (N + PP_TILE - 1) / PP_TILE;
    const float4* src = pm4;
    if (N % PP_TILE) {
        const uint32_t npad = tiles * PP_TILE;
        if (c->pp_src.reserve((size_t)npad * 16, fn)) return SFE_FAILURE;
        pp_pad_kernel<<<(npad + 255) / 256, 256, 0, s>>>(pm4, c->pp_src.as<float4>(), N, npad);
        if (check_launch(fn)) return SFE_FAILURE;
        src = c->pp_src.as<float4>();
    }
    const uint32_t per_cta = PP_THREADS * 2 * PP_T;
    const uint32_t iblocks = (ni + per_cta - 1) / per_cta;
    // choose the j-split count so that the grid is close to a whole number of waves of resident CTAs
    const uint32_t wave = (uint32_t)occ * (uint32_t)c->sm_count;
    const uint32_t max_splits = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(64, ((uint64_t)256 << 20) / ((uint64_t)ni * 16)));
    uint32_t splits = 1;
    if (iblocks < 8u * wave) {
        uint32_t best = 1; double best_eff = 0.0;
        for (uint32_t sp = 1; sp <= max_splits && sp <= tiles; sp++) {
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

    // equal-mass detection on the device; both specialisations are enqueued and the flag picks one
    if (c->pp_flag.reserve(sizeof(uint32_t), fn)) return SFE_FAILURE;
    uint32_t* flag = c->pp_flag.as<uint32_t>();
    SFE_CUDA(fn, cudaMemsetAsync(flag, 0, sizeof(uint32_t), s));
    pp_mass_scan_kernel<<<std::min<uint32_t>((N + 255) / 256, 4u * (uint32_t)c->sm_count), 256, 0, s>>>(pm4, N, flag);
    if (check_launch(fn)) return SFE_FAILURE;

    dim3 grid(iblocks, splits);
    kernel_um<<<grid, PP_THREADS, smem, s>>>(src, pm4, N, i0, ni, tiles, tps, eps2, c->pp_partial.as<float4>(), flag);
    if (check_launch(fn)) return SFE_FAILURE;
    kernel<<<grid, PP_THREADS, smem, s>>>(src, pm4, N, i0, ni, tiles, tps, eps2, c->pp_partial.as<float4>(), flag);
    if (check_launch(fn)) return SFE_FAILURE;
    pp_finish_kernel<<<(ni + 255) / 256, 256, 0, s>>>(c->pp_partial.as<float4>(), pm4, i0, ni, splits, G, acc, pot, flag);
    return check_launch(fn);
}

}  // namespace sfe
