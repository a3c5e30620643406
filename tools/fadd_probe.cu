// Dependent-add chain probes: how fast can ONE lane run the reference's fp32 left fold (tree.c:166-169)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fadd_probe tools/fadd_probe.cu && tools/fadd_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain_regs(float* out, long long* cyc, int n) {            // operands in registers
    float s = out[0], a = out[1], b = out[2], c = out[3], d = out[4];
    long long t0 = clock64();
    for (int i = 0; i < n; i += 4) { s = __fadd_rn(s, a); s = __fadd_rn(s, b); s = __fadd_rn(s, c); s = __fadd_rn(s, d); }
    long long t1 = clock64();
    out[0] = s; cyc[0] = t1 - t0;
}
__global__ void chain_lds(float* out, long long* cyc, int n) {             // operands from shared memory, LDS.128 per four adds
    __shared__ __align__(16) float buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = out[i & 7] * 1e-3f;
    __syncthreads();
    if (threadIdx.x != 0) return;
    float s = out[0];
    long long t0 = clock64();
    for (int r = 0; r < n / 1024; r++) {
        const float4* row = reinterpret_cast<const float4*>(buf);
#pragma unroll 8
        for (int v4 = 0; v4 < 256; v4++) { const float4 v = row[v4]; s = __fadd_rn(s, v.x); s = __fadd_rn(s, v.y); s = __fadd_rn(s, v.z); s = __fadd_rn(s, v.w); }
    }
    long long t1 = clock64();
    out[0] = s; cyc[0] = t1 - t0;
}
int main() {
    float* d; long long* c; cudaMalloc(&d, 64); cudaMalloc(&c, 8);
    float h[8] = {1.f, 1e-3f, 2e-3f, 3e-3f, 4e-3f, 1.f, 2.f, 3.f}; long long hc;
    const int n = 1 << 20;
    for (int threads : {1, 32}) {
        cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice);
        chain_regs<<<1, threads>>>(d, c, n); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
        printf("registers, %2d threads: %.2f cycles per dependent FADD\n", threads, (double)hc / n);
        cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice);
        chain_lds<<<1, threads>>>(d, c, n); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
        printf("LDS.128 + 4 FADD, %2d threads (lane 0 adds): %.2f cycles per member\n", threads, (double)hc / n);
    }
    return 0;
}
