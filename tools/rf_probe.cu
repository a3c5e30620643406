// rf_probe.cu - does register-file operand bandwidth limit packed FFMA2 on sm_100a?
// Times FFMA2 streams with 1, 2 or 3 fresh (non-reusable) 64-bit register operands, and the PP interaction
// body with its sources held in registers (no LDS), with and without the MUFU.RSQ.
// Output: achieved FMA-pipe lane-ops relative to 128 lanes/clk/SM at the SM clock seen by cudaEvents.
#include <cuda_runtime.h>
#include <stdio.h>
#define ITERS 2048
#define NA 16

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
    float2 a[NA], b = make_float2(seed, seed * 0.5f), c = make_float2(0.999f, 1.001f);
#pragma unroll
    for (int i = 0; i < NA; i++) a[i] = make_float2(seed + i, seed - i);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NA; i++) {
            if (MODE == 0) a[i] = __ffma2_rn(a[i], c, b);                                   // 1 fresh pair
            if (MODE == 1) a[i] = __ffma2_rn(a[i], a[(i + 1) % NA], b);                     // 2 fresh pairs
            if (MODE == 2) a[i] = __ffma2_rn(a[(i + 1) % NA], a[(i + 2) % NA], a[i]);       // 3 fresh pairs
            if (MODE == 3) a[i] = __ffma2_rn(a[(i + 1) % NA], a[(i + 1) % NA], a[i]);       // 2 fresh, one used twice
            if (MODE == 4) a[i] = __fadd2_rn(a[i], a[(i + 1) % NA]);                        // FADD2 2 fresh
            if (MODE == 5) a[i] = __fmul2_rn(a[(i + 3) % NA], a[(i + 1) % NA]);             // FMUL2 2 fresh
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NA; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// PP body, T targets, sources in registers (rotating through 4 pair records), MUFU optional
template <int T, bool MUFU>
__global__ void __launch_bounds__(256, 2) body(float* out, float seed) {
    float2 nx[T], ny[T], nz[T], ax[T], ay[T], az[T], ap[T];
    for (int k = 0; k < T; k++) {
        nx[k] = make_float2(seed * k, seed * k); ny[k] = make_float2(seed + k, seed + k); nz[k] = make_float2(seed - k, seed - k);
        ax[k] = ay[k] = az[k] = ap[k] = make_float2(0.f, 0.f);
    }
    float2 xj = make_float2(seed, 1.f - seed), yj = make_float2(0.3f, seed), zj = make_float2(seed, 0.7f), mj = make_float2(1e-6f, 2e-6f);
    const float2 e2 = make_float2(1e-6f, 1e-6f);
    for (int it = 0; it < ITERS * 2; it++) {
#pragma unroll
        for (int p = 0; p < 4; p++) {
            xj.x += 1e-3f; yj.y -= 1e-3f;   // keep the sources changing (2 scalar ops per 4*T*13 packed ops)
#pragma unroll
            for (int k = 0; k < T; k++) {
                const float2 dx = __fadd2_rn(xj, nx[k]);
                const float2 dy = __fadd2_rn(yj, ny[k]);
                const float2 dz = __fadd2_rn(zj, nz[k]);
                float2 r2 = __ffma2_rn(dx, dx, e2);
                r2 = __ffma2_rn(dy, dy, r2);
                r2 = __ffma2_rn(dz, dz, r2);
                float2 inv;
                if (MUFU) {
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(r2.x));
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(r2.y));
                } else {
                    inv = r2;   // no MUFU: 13 packed ops only (values are garbage, timing is what matters)
                }
                const float2 u = __fmul2_rn(mj, inv);
                const float2 i2 = __fmul2_rn(inv, inv);
                const float2 w = __fmul2_rn(u, i2);
                ax[k] = __ffma2_rn(w, dx, ax[k]);
                ay[k] = __ffma2_rn(w, dy, ay[k]);
                az[k] = __ffma2_rn(w, dz, az[k]);
                ap[k] = __fadd2_rn(ap[k], u);
            }
        }
    }
    float s = 0;
    for (int k = 0; k < T; k++) s += ax[k].x + ax[k].y + ay[k].x + ay[k].y + az[k].x + az[k].y + ap[k].x + ap[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static int sms;
static double time_ms(void (*launch)(float*), float* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) launch(out);   // warm the clocks up
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); launch(out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}
template <int MODE> void lk(float* out) { k<MODE><<<sms * 8, 256>>>(out, 1.0f); }
template <int T, bool M> void lb(float* out) { body<T, M><<<sms * 2, 256>>>(out, 0.5f); }

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    const double lanes_per_s = 128.0 * sms * 1.965e9;
    const char* names[] = {"ffma2 1 fresh pair", "ffma2 2 fresh pairs", "ffma2 3 fresh pairs", "ffma2 1 pair twice + 1", "fadd2 2 fresh pairs", "fmul2 2 fresh pairs"};
    void (*ks[])(float*) = {lk<0>, lk<1>, lk<2>, lk<3>, lk<4>, lk<5>};
    for (int m = 0; m < 6; m++) {
        double ms = time_ms(ks[m], out);
        double laneops = (double)sms * 8 * 256 * ITERS * NA * 2;
        printf("%-28s %8.3f ms  FMA-pipe utilisation %.3f (of 128 lanes/clk/SM at 1965 MHz)\n", names[m], ms, laneops / (ms * 1e-3) / lanes_per_s);
    }
    struct { const char* n; void (*f)(float*); int T; } bs[] = {
        {"body T4 with MUFU", lb<4, true>, 4}, {"body T4 no MUFU", lb<4, false>, 4},
        {"body T2 with MUFU", lb<2, true>, 2}, {"body T2 no MUFU", lb<2, false>, 2}};
    for (auto& b : bs) {
        double ms = time_ms(b.f, out);
        double laneops = (double)sms * 2 * 256 * (ITERS * 2.0) * 4 * b.T * 13 * 2;
        printf("%-28s %8.3f ms  FMA-pipe utilisation %.3f  -> %.2f TFLOP/s by the 20-flop count\n", b.n, ms, laneops / (ms * 1e-3) / lanes_per_s,
               laneops / 13.0 * 20.0 / (ms * 1e-3) / 1e12);
    }
    return 0;
}
