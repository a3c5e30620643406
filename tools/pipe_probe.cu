// pipe_probe.cu - per-instruction throughput probes for the PP inner loop on sm_100a.
// Reports warp-instructions per cycle per SM for several instruction mixes, using clock64 on device and
// the SM clock sampled from nvidia-smi is not needed: cycles are counted on-chip.
#include <cuda_runtime.h>
#include <stdio.h>
#define ITERS 2048
#define NA 8
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, float seed) {
    float2 a[NA], b = make_float2(seed, seed * 0.5f), c = make_float2(0.999f, 1.001f);
#pragma unroll
    for (int i = 0; i < NA; i++) a[i] = make_float2(seed + i, seed - i);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NA; i++) {
            if (MODE == 0) a[i] = __ffma2_rn(a[i], c, b);
            if (MODE == 1) a[i] = __fadd2_rn(a[i], b);
            if (MODE == 2) a[i] = __fmul2_rn(a[i], c);
            if (MODE == 3) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i].x)); }
            if (MODE == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i].x)); }
            if (MODE == 5) { a[i].x = fmaf(a[i].x, c.x, b.x); }
            if (MODE == 6) { // 13 packed : 2 mufu, independent
                float2 t = a[i];
#pragma unroll
                for (int r = 0; r < 13; r++) t = __ffma2_rn(t, c, b);
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(t.x));
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(t.y));
                a[i] = t;
            }
            if (MODE == 7) { // 13 packed : 1 mufu
                float2 t = a[i];
#pragma unroll
                for (int r = 0; r < 13; r++) t = __ffma2_rn(t, c, b);
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(t.x));
                a[i] = t;
            }
            if (MODE == 8) { // 6 ffma2 + 4 fadd2 + 3 fmul2, no mufu
                float2 t = a[i];
                t = __fadd2_rn(t, b); t = __fadd2_rn(t, c); t = __fadd2_rn(t, b); t = __fadd2_rn(t, c);
                t = __fmul2_rn(t, c); t = __fmul2_rn(t, c); t = __fmul2_rn(t, c);
#pragma unroll
                for (int r = 0; r < 6; r++) t = __ffma2_rn(t, c, b);
                a[i] = t;
            }
            if (MODE == 9) { // scalar: 26 ffma + 2 mufu
                float2 t = a[i];
#pragma unroll
                for (int r = 0; r < 13; r++) { t.x = fmaf(t.x, c.x, b.x); t.y = fmaf(t.y, c.y, b.y); }
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(t.x));
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(t.y));
                a[i] = t;
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NA; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int sms, float* out, long long* cyc, double instr_per_inner, int ctas_per_sm) {
    int grid = sms * ctas_per_sm;
    k<MODE><<<grid, 256>>>(out, cyc, 1.0f); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<grid, 256>>>(out, cyc, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[4096]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
    // warps per SM = ctas_per_sm * 8 ; instr per warp = ITERS*NA*instr_per_inner
    double winstr_per_sm = (double)ctas_per_sm * 8 * ITERS * NA * instr_per_inner;
    printf("%-28s ctas/sm %d : %.3f warp-instr/clk/SM  (%.2f clk per inner per SMSP-warp-slot)  kernel %.3f ms  eff clock %.0f MHz\n", name, ctas_per_sm,
           winstr_per_sm / avg, avg / (ITERS * NA) / (ctas_per_sm * 8 / 4.0), ms, avg / (ms * 1e-3) / 1e6);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    long long* cyc; cudaMalloc(&cyc, 4096 * 8);
    for (int c : {2, 8}) {
        run<0>("ffma2", p.multiProcessorCount, out, cyc, 1, c);
        run<1>("fadd2", p.multiProcessorCount, out, cyc, 1, c);
        run<2>("fmul2", p.multiProcessorCount, out, cyc, 1, c);
        run<3>("mufu.rsq", p.multiProcessorCount, out, cyc, 1, c);
        run<4>("mufu.ex2", p.multiProcessorCount, out, cyc, 1, c);
        run<5>("ffma", p.multiProcessorCount, out, cyc, 1, c);
        run<6>("13 ffma2 + 2 mufu", p.multiProcessorCount, out, cyc, 15, c);
        run<7>("13 ffma2 + 1 mufu", p.multiProcessorCount, out, cyc, 14, c);
        run<8>("6ffma2+4fadd2+3fmul2", p.multiProcessorCount, out, cyc, 13, c);
        run<9>("26 ffma + 2 mufu", p.multiProcessorCount, out, cyc, 28, c);
    }
    return 0;
}
