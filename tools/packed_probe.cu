// packed_probe.cu - do the packed fp32x2 instructions (FADD2 / FMUL2 / FFMA2) round each lane exactly like their scalar
// forms?  Compares bit patterns on random operands of several magnitudes, including denormal operands and results.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }
__device__ float make(uint32_t& s, int cls) {
    uint32_t m = rng(s) & 0x007FFFFFu, sign = rng(s) & 0x80000000u, e;
    uint32_t r = rng(s);
    if (cls == 0) e = 110u + (r % 30u);            // ~1e-5 .. 1e3
    else if (cls == 1) e = 1u + (r % 20u);         // tiny normals
    else if (cls == 2) e = 0u;                     // denormals
    else e = 60u + (r % 140u);                     // wide
    return __uint_as_float(sign | (e << 23) | m);
}
__global__ void probe(unsigned long long* bad, int cls) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u + cls;
    for (int it = 0; it < 4096; it++) {
        float a0 = make(s, cls), a1 = make(s, cls), b0 = make(s, cls == 2 ? 1 : cls), b1 = make(s, cls), c0 = make(s, cls), c1 = make(s, cls);
        float2 A = make_float2(a0, a1), B = make_float2(b0, b1), C = make_float2(c0, c1);
        float2 r = __fadd2_rn(A, B);
        if (__float_as_uint(r.x) != __float_as_uint(__fadd_rn(a0, b0)) || __float_as_uint(r.y) != __float_as_uint(__fadd_rn(a1, b1))) atomicAdd(&bad[0], 1ull);
        r = __fmul2_rn(A, B);
        if (__float_as_uint(r.x) != __float_as_uint(__fmul_rn(a0, b0)) || __float_as_uint(r.y) != __float_as_uint(__fmul_rn(a1, b1))) atomicAdd(&bad[1], 1ull);
        r = __ffma2_rn(A, B, C);
        if (__float_as_uint(r.x) != __float_as_uint(__fmaf_rn(a0, b0, c0)) || __float_as_uint(r.y) != __float_as_uint(__fmaf_rn(a1, b1, c1))) atomicAdd(&bad[2], 1ull);
        // near-cancellation (Kahan's z = t - s)
        float t0 = __fadd_rn(a0, b0), t1 = __fadd_rn(a1, b1);
        r = __fadd2_rn(make_float2(t0, t1), make_float2(-a0, -a1));
        if (__float_as_uint(r.x) != __float_as_uint(__fsub_rn(t0, a0)) || __float_as_uint(r.y) != __float_as_uint(__fsub_rn(t1, a1))) atomicAdd(&bad[3], 1ull);
    }
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 32);
    for (int cls = 0; cls < 4; cls++) {
        cudaMemset(d, 0, 32);
        probe<<<256, 256>>>(d, cls);
        unsigned long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("class %d: mismatches add2 %llu  mul2 %llu  fma2 %llu  cancel %llu  (of %d each)\n", cls, h[0], h[1], h[2], h[3], 256 * 256 * 4096);
    }
    return 0;
}
