// fp32_peak.cu - measured non-tensor FP32 peaks on the box (SURVEY 8d asks for a measured FMA-loop peak
// next to the 2*128*SMs*clock figure).  Prints one JSON line.
//   ffma    : dependent-chain-free scalar FFMA loop
//   ffma2   : packed fma.rn.f32x2 loop (sm_100+)
//   mix     : the PP inner-loop mix (4 FADD2 + 6 FFMA2 + 3 FMUL2 + 2 MUFU.RSQ per pair)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
    float2 a[8], b = make_float2(seed, seed * 0.5f), c = make_float2(0.999f, 1.001f);
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(seed + i, seed - i);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, c.x, b.x); a[i].y = fmaf(a[i].y, c.y, b.y); }
            if (MODE == 1) { a[i] = __ffma2_rn(a[i], c, b); }
            if (MODE == 2) {  // PP mix on one accumulator set
                float2 dx = __fadd2_rn(a[i], b), dy = __fadd2_rn(a[i], c), dz = __fadd2_rn(b, a[i]);
                float2 r2 = __ffma2_rn(dx, dx, c); r2 = __ffma2_rn(dy, dy, r2); r2 = __ffma2_rn(dz, dz, r2);
                float2 inv;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(r2.x));
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(r2.y));
                float2 u = __fmul2_rn(b, inv), i2 = __fmul2_rn(inv, inv), w = __fmul2_rn(u, i2);
                a[i] = __ffma2_rn(w, dx, a[i]); a[i] = __ffma2_rn(w, dy, a[i]); a[i] = __ffma2_rn(w, dz, a[i]);
                a[i] = __fadd2_rn(a[i], u);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> double run(int sms, float* out, double flop_per_inner) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = sms * 8;
    k<MODE><<<grid, 256>>>(out, 1.0f); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k<MODE><<<grid, 256>>>(out, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flop = (double)grid * 256 * ITERS * 8 * flop_per_inner;
    return flop / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    double t0 = run<0>(p.multiProcessorCount, out, 4.0);
    double t1 = run<1>(p.multiProcessorCount, out, 4.0);
    // mode 2: per inner = 2 interactions = 40 flop by the classic count (13 FMA-pipe lane-ops + 1 rsqrt each)
    double t2 = run<2>(p.multiProcessorCount, out, 40.0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"nominal_fp32_tflops\": %.2f, "
           "\"ffma_tflops\": %.2f, \"ffma2_tflops\": %.2f, \"pp_mix_tflops_20flop\": %.2f}\n",
           p.name, p.multiProcessorCount, khz, 2.0 * 128 * p.multiProcessorCount * khz * 1e3 / 1e12, t0, t1, t2);
    return 0;
}
