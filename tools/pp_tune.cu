// pp_tune.cu - times variants of the PP kernel template (starflood_b200/csrc/sfe_pp_kernel.cuh) on the box.
// Scratch tool for choosing T / CTA size / occupancy / tile size; prints one line per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/pp_tune tools/pp_tune.cu
#include "../starflood_b200/csrc/sfe_pp_kernel.cuh"
#include <vector>
#include <cmath>
#include <cstdlib>

using namespace sfe;

__global__ void pair_kernel(const float4* __restrict__ pm4, float4* __restrict__ pairs, uint32_t N, uint32_t npairs) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const float4 pad = make_float4(PP_PAD_POS, PP_PAD_POS, PP_PAD_POS, 0.0f);
    float4 a = (2u * p < N) ? pm4[2u * p] : pad;
    float4 b = (2u * p + 1u < N) ? pm4[2u * p + 1u] : pad;
    pairs[2u * p] = make_float4(a.x, b.x, a.y, b.y);
    pairs[2u * p + 1u] = make_float4(a.z, b.z, a.w, b.w);
}

static int g_sms = 148;
static std::vector<float4> g_ref;

template <int T, int THREADS, int MINB, int TILE, bool ST, int UNR>
void run(const char* name, const float4* pm4, float4* pairs, float4* partial, uint32_t N) {
    const uint32_t tiles = (N + TILE - 1) / TILE;
    const uint32_t npairs = tiles * (TILE / 2);
    pair_kernel<<<(npairs + 255) / 256, 256>>>(pm4, pairs, N, npairs);
    const uint32_t iblocks = (N + THREADS * T - 1) / (THREADS * T);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pp_kernel<T, THREADS, MINB, TILE, ST, UNR>, THREADS, 0);
    const uint32_t wave = (uint32_t)(occ * g_sms);
    // j-splits so that the grid is a near-whole number of waves
    uint32_t best = 1; double best_eff = 0;
    for (uint32_t sp = 1; sp <= 64 && sp <= tiles; sp++) {
        uint32_t tps = (tiles + sp - 1) / sp, rsp = (tiles + tps - 1) / tps;
        uint64_t ctas = (uint64_t)iblocks * rsp, waves = (ctas + wave - 1) / wave;
        double eff = (double)iblocks * tiles / ((double)waves * wave * tps);
        if (tps < 8) eff *= 0.9;
        if (eff > best_eff + 1e-9) { best_eff = eff; best = rsp; }
    }
    uint32_t tps = (tiles + best - 1) / best, splits = (tiles + tps - 1) / tps;
    dim3 grid(iblocks, splits);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, pp_kernel<T, THREADS, MINB, TILE, ST, UNR>);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best_ms = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        pp_kernel<T, THREADS, MINB, TILE, ST, UNR><<<grid, THREADS>>>(pairs, pm4, N, 0, N, tiles, tps, 1e-6f, partial);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best_ms) best_ms = ms;
    }
    cudaError_t err = cudaGetLastError();
    // combine splits on the host for a cross-check against the first variant
    std::vector<float4> h((size_t)splits * N);
    cudaMemcpy(h.data(), partial, h.size() * sizeof(float4), cudaMemcpyDeviceToHost);
    std::vector<float4> out(N);
    for (uint32_t i = 0; i < N; i++) {
        double x = 0, y = 0, z = 0, w = 0;
        for (uint32_t s = 0; s < splits; s++) { float4 v = h[(size_t)s * N + i]; x += v.x; y += v.y; z += v.z; w += v.w; }
        out[i] = make_float4((float)x, (float)y, (float)z, (float)w);
    }
    double num = 0, den = 0;
    if (g_ref.empty()) g_ref = out;
    for (uint32_t i = 0; i < N; i++) {
        double dx = out[i].x - g_ref[i].x, dy = out[i].y - g_ref[i].y, dz = out[i].z - g_ref[i].z;
        num += dx * dx + dy * dy + dz * dz;
        den += (double)g_ref[i].x * g_ref[i].x + (double)g_ref[i].y * g_ref[i].y + (double)g_ref[i].z * g_ref[i].z;
    }
    double tf = 20.0 * (double)N * N / (best_ms * 1e-3) / 1e12;
    printf("%-34s regs %3d occ %d grid %ux%u eff %.3f  %8.3f ms  %6.2f TFLOP/s  relrms-vs-first %.2e  %s\n", name, fa.numRegs, occ,
           iblocks, splits, best_eff, best_ms, tf, sqrt(num / den), err == cudaSuccess ? "" : cudaGetErrorString(err));
    fflush(stdout);
}

__global__ void pad_kernel(const float4* __restrict__ pm4, float4* __restrict__ src, uint32_t N, uint32_t npad) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < npad) src[j] = j < N ? pm4[j] : make_float4(PP_PAD_POS, PP_PAD_POS, PP_PAD_POS, 0.f);
}

template <int T, int THREADS, int MINB, int TILE, int UNR, int ACC = 0, bool UM = false>
void run_tp(const char* name, const float4* pm4, float4* srcbuf, float4* partial, uint32_t N) {
    const uint32_t tiles = (N + TILE - 1) / TILE, npad = tiles * TILE;
    pad_kernel<<<(npad + 255) / 256, 256>>>(pm4, srcbuf, N, npad);
    const uint32_t per_cta = THREADS * 2 * T;
    const uint32_t iblocks = (N + per_cta - 1) / per_cta;
    int occ = 0;
    const size_t smem = pp_tp_smem_bytes<T, THREADS, TILE>();
    cudaFuncSetAttribute(pp_kernel_tp<T, THREADS, MINB, TILE, UNR, ACC, UM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pp_kernel_tp<T, THREADS, MINB, TILE, UNR, ACC, UM>, THREADS, smem);
    const uint32_t wave = (uint32_t)(occ * g_sms);
    uint32_t best = 1; double best_eff = 0;
    for (uint32_t sp = 1; sp <= 64 && sp <= tiles; sp++) {
        uint32_t tps = (tiles + sp - 1) / sp, rsp = (tiles + tps - 1) / tps;
        uint64_t ctas = (uint64_t)iblocks * rsp, waves = (ctas + wave - 1) / wave;
        double eff = (double)iblocks * tiles / ((double)waves * wave * tps);
        if (tps < 8) eff *= 0.9;
        if (eff > best_eff + 1e-9) { best_eff = eff; best = rsp; }
    }
    uint32_t tps = (tiles + best - 1) / best, splits = (tiles + tps - 1) / tps;
    dim3 grid(iblocks, splits);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, pp_kernel_tp<T, THREADS, MINB, TILE, UNR, ACC, UM>);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best_ms = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        pp_kernel_tp<T, THREADS, MINB, TILE, UNR, ACC, UM><<<grid, THREADS, smem>>>(srcbuf, pm4, N, 0, N, tiles, tps, 1e-6f, partial);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best_ms) best_ms = ms;
    }
    cudaError_t err = cudaGetLastError();
    std::vector<float4> h((size_t)splits * N);
    cudaMemcpy(h.data(), partial, h.size() * sizeof(float4), cudaMemcpyDeviceToHost);
    double num = 0, den = 0;
    for (uint32_t i = 0; i < N; i++) {
        double x = 0, y = 0, z = 0;
        for (uint32_t s = 0; s < splits; s++) { float4 v = h[(size_t)s * N + i]; x += v.x; y += v.y; z += v.z; }
        double dx = (float)x - g_ref[i].x, dy = (float)y - g_ref[i].y, dz = (float)z - g_ref[i].z;
        num += dx * dx + dy * dy + dz * dz;
        den += (double)g_ref[i].x * g_ref[i].x + (double)g_ref[i].y * g_ref[i].y + (double)g_ref[i].z * g_ref[i].z;
    }
    double tf = 20.0 * (double)N * N / (best_ms * 1e-3) / 1e12;
    printf("%-34s regs %3d occ %d grid %ux%u eff %.3f  %8.3f ms  %6.2f TFLOP/s  relrms-vs-first %.2e  %s\n", name, fa.numRegs, occ,
           iblocks, splits, best_eff, best_ms, tf, sqrt(num / den), err == cudaSuccess ? "" : cudaGetErrorString(err));
    fflush(stdout);
}

int main(int argc, char** argv) {
    uint32_t N = argc > 1 ? (uint32_t)atoi(argv[1]) : 262144u;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    g_sms = p.multiProcessorCount;
    std::vector<float4> h(N);
    srand(1);
    for (uint32_t i = 0; i < N; i++) {
        float x, y, z;
        do { x = 2.f * rand() / RAND_MAX - 1.f; y = 2.f * rand() / RAND_MAX - 1.f; z = 2.f * rand() / RAND_MAX - 1.f; } while (x * x + y * y + z * z > 1.f);
        h[i] = make_float4(x, y, z, 4.18879f / N);
    }
    float4 *pm4, *pairs, *partial;
    cudaMalloc(&pm4, (size_t)N * 16);
    cudaMalloc(&pairs, ((size_t)N + 4096) * 16);
    cudaMalloc(&partial, (size_t)N * 64 * 16);
    cudaMemcpy(pm4, h.data(), (size_t)N * 16, cudaMemcpyHostToDevice);
    printf("N = %u, %s, %d SMs; nominal FP32 peak %.2f TFLOP/s\n", N, p.name, g_sms, 2.0 * 128 * g_sms * 1.965e9 / 1e12);
    run<4, 256, 2, 512, false, 4>("T4 256x2 tile512 reg-tot u4 (r01)", pm4, pairs, partial, N);
    //      T(pairs) THR MINB TILE UNROLL ACC UM
    run_tp<6, 128, 2, 512, 2>("TP T6 128x2 u2 packed (current)", pm4, pairs, partial, N);
    run_tp<6, 128, 2, 512, 2, 1>("TP T6 128x2 u2 scalar-acc", pm4, pairs, partial, N);
    run_tp<4, 128, 3, 512, 4, 1>("TP T4 128x3 u4 scalar-acc", pm4, pairs, partial, N);
    run_tp<4, 128, 4, 512, 4, 1>("TP T4 128x4 u4 scalar-acc", pm4, pairs, partial, N);
    run_tp<4, 256, 2, 512, 4, 1>("TP T4 256x2 u4 scalar-acc", pm4, pairs, partial, N);
    run_tp<3, 128, 5, 512, 4, 1>("TP T3 128x5 u4 scalar-acc", pm4, pairs, partial, N);
    run_tp<8, 64, 4, 512, 2, 1>("TP T8 64x4 u2 scalar-acc", pm4, pairs, partial, N);
    run_tp<6, 128, 2, 512, 2, 0, true>("TP T6 128x2 u2 packed UM", pm4, pairs, partial, N);
    run_tp<6, 128, 2, 512, 2, 1, true>("TP T6 128x2 u2 scalar-acc UM", pm4, pairs, partial, N);
    run_tp<4, 128, 4, 512, 4, 1, true>("TP T4 128x4 u4 scalar-acc UM", pm4, pairs, partial, N);
    return 0;
}
