// sfe_vis.cu - vis_draw() on the device (visualization.c:253-739; SURVEY 8f row N4, second half).
//
// The reference renders every frame by splatting SPATIAL_SAMPLES (128) Gaussian-jittered samples per particle into an
// INTEGER accumulation buffer with atomic adds (visualization.c:503-683) and then maps the buffer to floats
// (:691-735).  On the host that loop costs as much as the Barnes-Hut solver it visualises (N = 2^20: 1.3e8 samples,
// each a pcg4d hash, logf, sqrtf, cosf, sinf); once the solver runs in 3 ms it would be 99 % of a step.
//
// Parity: integer adds commute, so the accumulation buffer is a pure function of the inputs and is reproduced BIT FOR BIT.
// That needs the host libm's logf / sinf / cosf bit patterns (a correctly rounded cosf differs from glibc's in ~1 % of
// the arguments the splat draws, enough to move a sample across a pixel edge now and then): sfe_libm.cuh restates glibc 2.39's
// algorithms in IEEE double operations and is pinned exhaustively against the host library (tools/libm_probe.cpp).  All
// fp32 arithmetic is written without contraction (__fmul_rn / __fadd_rn) in the reference's operation order.
//
// Kernels
//   vis_mvp_kernel     matMVP = matP * (matV * matM) with ortho_scale *= (real)(1.0 / average_radius)   (:418-489)
//   vis_splat_kernel   one warp per particle, one lane per sample; the samples of a particle fall within a few pixels of
//                      each other, so they are first counted in an 8x8-pixel window in shared memory and only the
//                      non-empty bins go to global memory as {count*r, count*g, count*b, count} (every sample of a
//                      particle carries the same colour): ~20 bins instead of 128 samples, 6x fewer global atomics
//   vis_finish_kernel  pixel_value_scale * accum, tanh tonemap in double, alpha = 1                      (:691-735)
// Bound: instruction issue (ncu: issue slots 75 % busy, 28.5 of 32 lanes active, DRAM 27 MB per frame): per sample a pcg4d
// hash, ~35 double-precision operations for logf / sinf / cosf and a dozen int<->float<->double conversions; global traffic
// is the 16 bytes per particle read plus ~20 x 4 atomics per particle.  [B200] N = 2^20, 128 samples, 1080 x 1080:
// 0.92 ms per frame against 331 ms for the reference on 16 host threads.
#include "sfe_internal.cuh"
#include "sfe_libm.cuh"
#include <stdlib.h>

namespace sfe {

constexpr int VIS_THREADS = 256;
constexpr int VIS_WARPS = VIS_THREADS / 32;
#ifndef SFE_VIS_MINB
#define SFE_VIS_MINB 5                           // resident CTAs per SM the splat kernel is compiled for (46 registers)
#endif
constexpr int VIS_WIN = 8;                       // window edge in pixels (sigma = 0.75 px: +-4 px is 5.3 sigma)

struct VisInfo { float mvp[16]; double rbar; };  // device block, copied back to the host by the entry points

__device__ __forceinline__ void pcg4d_vis(uint32_t (&v)[4]) {   // rng.h:13-40
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = v[k] * 0x0019660Du + 0x3C6EF35Fu;
    v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] ^= v[k] >> 16;
    v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
}

struct Mat16 { float m[16]; };

// visualization.c:418-489 with matV = I: matMVP = matP * (matV * matM); both products in the reference's term order
// (k = 3, 2, 1, 0), every product and sum rounded separately.
__global__ void vis_mvp_kernel(Mat16 vm, float ortho, int autoscale, const double* __restrict__ rbar, VisInfo* __restrict__ info) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double r = *rbar;
    if (autoscale) ortho = __fmul_rn(ortho, (float)__ddiv_rn(1.0, r));          // ortho_scale *= (real)(1.0 / average_radius)
    float P[16];
    for (int i = 0; i < 16; i++) P[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) P[4 * i + j] = __fmul_rn(P[4 * i + j], ortho);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float c = __fadd_rn(__fmul_rn(P[4 * i + 3], vm.m[12 + j]), __fmul_rn(P[4 * i + 2], vm.m[8 + j]));
            c = __fadd_rn(c, __fmul_rn(P[4 * i + 1], vm.m[4 + j]));
            c = __fadd_rn(c, __fmul_rn(P[4 * i + 0], vm.m[j]));
            info->mvp[4 * i + j] = c;
        }
    info->rbar = r;
}

__device__ __forceinline__ void splat_add(int32_t* __restrict__ ab, uint32_t pix, int c0, int c1, int c2, int n) {
    int32_t* q = ab + 4u * (size_t)pix;
    atomicAdd(q + 0, n * c0); atomicAdd(q + 1, n * c1); atomicAdd(q + 2, n * c2); atomicAdd(q + 3, n);
}

__global__ void __launch_bounds__(VIS_THREADS, SFE_VIS_MINB) vis_splat_kernel(const float4* __restrict__ pm4, const float* __restrict__ pot,
                                                                uint32_t i0, uint32_t i1, uint32_t step_number,
                                                                const VisInfo* __restrict__ info, uint32_t sizex, uint32_t sizey,
                                                                uint32_t samples, int32_t* __restrict__ ab) {
    __shared__ uint32_t s_win[VIS_WARPS][VIS_WIN * VIS_WIN];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t* win = s_win[warp];
    win[lane] = 0u; win[lane + 32u] = 0u;
    __syncwarp();
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; k++) m[k] = info->mvp[k];              // rows 0 and 1 are all the splat needs
    const float TAUF = 6.283185307179586f;
    const float hy = __fmul_rn(0.5f, (float)sizey);               // (real)0.5 * (real)sizey
    const float ox = __fmul_rn(0.5f, (float)sizex), oy = hy;
    const uint32_t nwarps = gridDim.x * VIS_WARPS;
    for (uint32_t idx = i0 + blockIdx.x * VIS_WARPS + warp; idx < i1; idx += nwarps) {      // particles [i0, i1): a rank's slice, or all
        const float4 p = __ldg(&pm4[idx]);
        // y = A * x with x[3] = 1 (visualization.c:536-552): ((a0 x0 + a1 x1) + a2 x2) + a3 x3
        const float px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], p.x), __fmul_rn(m[1], p.y)), __fmul_rn(m[2], p.z)), __fmul_rn(m[3], 1.0f));
        const float py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], p.x), __fmul_rn(m[5], p.y)), __fmul_rn(m[6], p.z)), __fmul_rn(m[7], 1.0f));
        // palette (visualization.c:554-560, :563-568): lanes 0..2 evaluate one channel each
        int col = 0;
        if (lane < 3u) {
            const float phase = __fmul_rn(1.000e-1f, __ldg(&pot[idx]));
            const float ck = lane == 0u ? (float)(6.283185307179586 * (0.0 / 3.0))
                           : lane == 1u ? (float)(6.283185307179586 * (1.0 / 3.0)) : (float)(6.283185307179586 * (2.0 / 3.0));
            const float c = __fadd_rn(__fmul_rn(0.5f, sfe_libm::cosf_glibc(__fsub_rn(__fmul_rn(TAUF, phase), ck))), 0.5f);
            const float c100 = __fmul_rn(100.0f, c);
            col = (c100 > -2147483648.0f && c100 < 2147483648.0f) ? __float2int_rz(c100) : (int)0x80000000;   // cvttss2si
        }
        const int c0 = __shfl_sync(0xFFFFFFFFu, col, 0), c1 = __shfl_sync(0xFFFFFFFFu, col, 1), c2 = __shfl_sync(0xFFFFFFFFu, col, 2);
        const float tx = __fmul_rn(hy, px), ty = __fmul_rn(hy, py);                     // (0.5 * sizey) * p_i
        // window origin: the pixel of the un-jittered sample, minus half a window (any choice is correct; this one catches most)
        const float gx = __fadd_rn(__fadd_rn(-0.5f, tx), ox), gy = __fadd_rn(__fadd_rn(-0.5f, ty), oy);
        const bool centre_ok = fabsf(gx) < 1.0e9f && fabsf(gy) < 1.0e9f;
        const int wx0 = centre_ok ? (int)floorf(gx) - VIS_WIN / 2 : 0, wy0 = centre_ok ? (int)floorf(gy) - VIS_WIN / 2 : 0;
        for (uint32_t s0 = 0; s0 < samples; s0 += 32u) {
            const uint32_t samp = s0 + lane;
            if (samp < samples) {
                bool continue_sample = true;
                uint32_t h[4] = {0xB79ABC95u + samp, 0xE0DA3F84u + idx, 0xAB75F07Bu + step_number, 0xCF52CA01u};
                pcg4d_vis(h);
                const float INV = 2.3283062977608182109179324470460414886474609375e-10f;   // (real)INV_PCG32_MAX
                const float r2 = __fmul_rn(INV, (float)h[2]), r3 = __fmul_rn(INV, (float)h[3]);
                // Box-Muller (visualization.c:583-586).  r2 = 0 gives logf = -inf, an infinite radius and a coordinate that
                // is +-inf or NaN: the reference drops such a sample at the bounds test, here it is dropped at once; any
                // other r2 is a normal float below 1 and ang lies in [0, tau], the ranges the branch-light variants cover
                if (r2 == 0.0f) continue_sample = false;
                const float rad = __fsqrt_rn(__fmul_rn(-2.0f, sfe_libm::logf_glibc_below_one(continue_sample ? r2 : 0.5f)));
                const float ang = __fmul_rn(TAUF, r3);
                float sn, cs;
                sfe_libm::sincosf_glibc_fast(ang, &sn, &cs);     // cosf(ang), sinf(ang) of the host libm, evaluated together
                const float n0 = __fmul_rn(rad, cs), n1 = __fmul_rn(rad, sn);
                // coord = (int)(samp_offset + -0.5 + 0.5*sizey*p + 0.5*size)  (visualization.c:660-663), left to right
                const float fx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(0.750f, n0), -0.5f), tx), ox);
                const float fy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(0.750f, n1), -0.5f), ty), oy);
                // (int) of NaN / out-of-range is INT_MIN on x86-64 and fails the bounds test
                if (continue_sample && fx > -2147483648.0f && fx < 2147483648.0f && fy > -2147483648.0f && fy < 2147483648.0f) {
                    const int cx = __float2int_rz(fx), cy = __float2int_rz(fy);
                    if (0 <= cx && cx < (int)sizex && 0 <= cy && cy < (int)sizey) {
                        const uint32_t wx = (uint32_t)(cx - wx0), wy = (uint32_t)(cy - wy0);
                        if (wx < (uint32_t)VIS_WIN && wy < (uint32_t)VIS_WIN) atomicAdd(&win[wy * VIS_WIN + wx], 1u);
                        else splat_add(ab, sizex * (uint32_t)cy + (uint32_t)cx, c0, c1, c2, 1);
                    }
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (uint32_t b = lane; b < (uint32_t)(VIS_WIN * VIS_WIN); b += 32u) {
            const uint32_t n = win[b];
            if (n) {
                win[b] = 0u;
                const uint32_t cx = (uint32_t)(wx0 + (int)(b % VIS_WIN)), cy = (uint32_t)(wy0 + (int)(b / VIS_WIN));
                splat_add(ab, sizex * cy + cx, c0, c1, c2, (int)n);
            }
        }
        __syncwarp();
    }
}

// visualization.c:699-735
__global__ void __launch_bounds__(256) vis_finish_kernel(const int32_t* __restrict__ ab, uint32_t npix, double scale, int tonemap,
                                                         float* __restrict__ rb) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        const int4 a = reinterpret_cast<const int4*>(ab)[i];
        double v[3] = {__dmul_rn(scale, (double)a.x), __dmul_rn(scale, (double)a.y), __dmul_rn(scale, (double)a.z)};
        if (tonemap) { v[0] = tanh(v[0]); v[1] = tanh(v[1]); v[2] = tanh(v[2]); }
        reinterpret_cast<float4*>(rb)[i] = make_float4((float)v[0], (float)v[1], (float)v[2], 1.0f);
    }
}

__global__ void vis_pack_kernel(const float* __restrict__ pos, float4* __restrict__ pm4, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pm4[i] = make_float4(pos[3u * i], pos[3u * i + 1u], pos[3u * i + 2u], 0.f);
}

static uint32_t vis_ctas_per_sm() {      // grid size of the splat in CTAs per SM; SFE_VIS_CTAS_PER_SM overrides (tuning knob)
    static uint32_t v = 0;
    if (!v) { const char* e = getenv("SFE_VIS_CTAS_PER_SM"); v = e && atoi(e) > 0 ? (uint32_t)atoi(e) : 24u; }
    return v;
}

// Enqueue the whole frame on stream s: mean radius -> MVP -> clear -> splat -> finish.  ab / rb / info are device buffers.
// sums the accumulation buffers of all ranks (their windows, at byte offset `off`) into this rank's own buffer: integer
// adds commute, so the result is the single-rank buffer bit for bit
__global__ void vis_reduce_kernel(PeerPtrs pp, uint64_t off, uint32_t words, int32_t* __restrict__ out) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < words / 4u; k += gridDim.x * blockDim.x) {
        int4 acc = make_int4(0, 0, 0, 0);
        for (uint32_t r = 0; r < pp.world; r++) {
            const int4 v = reinterpret_cast<const int4*>(pp.base[r] + off)[k];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<int4*>(out)[k] = acc;
    }
}

// `slice`: multi-rank mode - this rank splats the particles [i0, i1) into its window at vis_off, the ranks exchange arrival
// flags, every rank sums all windows into `ab`, and a second exchange keeps anybody from clearing its window for the
// next frame while a peer still reads it.
struct VisSlice { bool on; uint32_t i0, i1; uint64_t vis_off; };

static int launch_vis(Ctx* c, const char* fn, const float4* pm4, const float* pot, uint32_t N, uint32_t step_number,
                      const float mat_vm[16], float ortho_scale, int autoscale, uint32_t sizex, uint32_t sizey,
                      uint32_t samples, double pixel_value_scale, int tonemap, int32_t* ab, float* rb, VisInfo* info,
                      cudaStream_t s, VisSlice slice = VisSlice{false, 0u, 0u, 0u}) {
    const double* rbar = nullptr;
    if (launch_mean_radius(c, pm4, N, &rbar, s)) return SFE_FAILURE;
    Mat16 vm; memcpy(vm.m, mat_vm, sizeof vm.m);
    vis_mvp_kernel<<<1, 32, 0, s>>>(vm, ortho_scale, autoscale, rbar, info);
    if (check_launch(fn)) return SFE_FAILURE;
    const size_t npix = (size_t)sizex * sizey;
    const uint32_t i0 = slice.on ? slice.i0 : 0u, i1 = slice.on ? slice.i1 : N;
    int32_t* target = slice.on ? reinterpret_cast<int32_t*>((char*)c->peer.local + slice.vis_off) : ab;
    SFE_CUDA(fn, cudaMemsetAsync(target, 0, npix * 16, s));                                // visualization.c:334-342
    const uint32_t want = (i1 - i0 + VIS_WARPS - 1) / VIS_WARPS, cap = (uint32_t)c->sm_count * vis_ctas_per_sm();
    vis_splat_kernel<<<want < cap ? (want ? want : 1u) : cap, VIS_THREADS, 0, s>>>(pm4, pot, i0, i1, step_number, info, sizex, sizey,
                                                                                 samples, target);
    if (check_launch(fn)) return SFE_FAILURE;
    if (slice.on) {
        if (launch_peer_barrier(c, s, fn)) return SFE_FAILURE;                             // every rank's slice has landed
        vis_reduce_kernel<<<c->sm_count * 8, 256, 0, s>>>(peer_ptrs(c), slice.vis_off, (uint32_t)(npix * 4), ab);
        if (check_launch(fn)) return SFE_FAILURE;
        if (launch_peer_barrier(c, s, fn)) return SFE_FAILURE;                             // ... and has been read by everybody
    }
    const uint32_t fb = (uint32_t)((npix + 255) / 256), fcap = (uint32_t)c->sm_count * 8u;
    vis_finish_kernel<<<fb < fcap ? (fb ? fb : 1u) : fcap, 256, 0, s>>>(ab, (uint32_t)npix, pixel_value_scale, tonemap, rb);
    return check_launch(fn);
}

static int vis_args_ok(const char* fn, uint32_t N, const float* mat_vm, uint32_t sizex, uint32_t sizey, const void* rb) {
    if (!mat_vm || !rb) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if (N == 0) { set_error(fn, "N must be at least 1 (the reference divides by N, visualization.c:304)"); return SFE_FAILURE; }
    if (sizex == 0 || sizey == 0 || (uint64_t)sizex * sizey > (1ull << 28)) {
        set_error(fn, "image of %u x %u pixels is not supported", sizex, sizey);
        return SFE_FAILURE;
    }
    return SFE_SUCCESS;
}

// copies the results of a frame back and waits for them
static int vis_fetch(Ctx* c, const char* fn, uint32_t sizex, uint32_t sizey, int32_t* atomic_buffer, float* render_buffer,
                     double* average_radius, float* mvp) {
    const size_t npix = (size_t)sizex * sizey;
    if (atomic_buffer && d2h(fn, atomic_buffer, c->vis_atomic.p, npix * 16, c->stream)) return SFE_FAILURE;
    if (d2h(fn, render_buffer, c->vis_render.p, npix * 16, c->stream)) return SFE_FAILURE;
    VisInfo hinfo;
    SFE_CUDA(fn, cudaMemcpyAsync(&hinfo, c->vis_info.p, sizeof hinfo, cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(fn, cudaStreamSynchronize(c->stream));
    if (average_radius) *average_radius = hinfo.rbar;
    if (mvp) memcpy(mvp, hinfo.mvp, sizeof hinfo.mvp);
    return SFE_SUCCESS;
}

}  // namespace sfe

using namespace sfe;

extern "C" int sfe_vis_draw(const float* pos, const float* pot, uint32_t N, uint32_t step_number, const float mat_vm[16],
                            float ortho_scale, int autoscale, uint32_t sizex, uint32_t sizey, uint32_t samples,
                            double pixel_value_scale, int tonemap, int32_t* atomic_buffer, float* render_buffer,
                            double* average_radius, float mvp[16]) {
    const char* fn = "sfe_vis_draw";
    if (ensure_init(fn)) return SFE_FAILURE;
    if (!pos || !pot) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if (vis_args_ok(fn, N, mat_vm, sizex, sizey, render_buffer)) return SFE_FAILURE;
    Ctx* c = ctx();
    const size_t npix = (size_t)sizex * sizey;
    if (c->d_pos.reserve((size_t)N * 12, fn) || c->d_pot.reserve((size_t)N * 4, fn) || c->vis_pm4.reserve((size_t)N * 16, fn) ||
        c->vis_atomic.reserve(npix * 16, fn) || c->vis_render.reserve(npix * 16, fn) || c->vis_info.reserve(sizeof(VisInfo), fn))
        return SFE_FAILURE;
    memset(&c->timings, 0, sizeof c->timings);
    c->launches = 0;
    cudaEventRecord(c->ev[0], c->stream);
    if (h2d(fn, c->d_pos.p, pos, (size_t)N * 12, c->stream)) return SFE_FAILURE;
    if (h2d(fn, c->d_pot.p, pot, (size_t)N * 4, c->stream)) return SFE_FAILURE;
    cudaEventRecord(c->ev[1], c->stream);
    vis_pack_kernel<<<(N + 255) / 256, 256, 0, c->stream>>>(c->d_pos.as<float>(), c->vis_pm4.as<float4>(), N);
    if (check_launch(fn)) return SFE_FAILURE;
    if (launch_vis(c, fn, c->vis_pm4.as<float4>(), c->d_pot.as<float>(), N, step_number, mat_vm, ortho_scale, autoscale, sizex,
                   sizey, samples, pixel_value_scale, tonemap, c->vis_atomic.as<int32_t>(), c->vis_render.as<float>(),
                   c->vis_info.as<VisInfo>(), c->stream))
        return SFE_FAILURE;
    cudaEventRecord(c->ev[3], c->stream);
    if (vis_fetch(c, fn, sizex, sizey, atomic_buffer, render_buffer, average_radius, mvp)) return SFE_FAILURE;
    cudaEventRecord(c->ev[4], c->stream);
    SFE_CUDA(fn, cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&c->timings.h2d_ms, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&c->timings.solve_ms, c->ev[1], c->ev[3]);
    cudaEventElapsedTime(&c->timings.d2h_ms, c->ev[3], c->ev[4]);
    cudaEventElapsedTime(&c->timings.total_ms, c->ev[0], c->ev[4]);
    c->timings.launches = c->launches;
    return SFE_SUCCESS;
}

extern "C" int sfe_sim_vis_draw(uint32_t step_number, const float mat_vm[16], float ortho_scale, int autoscale, uint32_t sizex,
                                uint32_t sizey, uint32_t samples, double pixel_value_scale, int tonemap,
                                int32_t* atomic_buffer, float* render_buffer, double* average_radius, float mvp[16]) {
    const char* fn = "sfe_sim_vis_draw";
    if (ensure_init(fn)) return SFE_FAILURE;
    Ctx* c = ctx();
    const float4* pm4; const float* pot; uint32_t N, i0 = 0u, ni = 0u; uint64_t vis_off = 0u; bool sliced = false;
    if (sim_resident_views(c, &pm4, &pot, &N, fn, &i0, &ni, &vis_off, &sliced)) return SFE_FAILURE;
    if (vis_args_ok(fn, N, mat_vm, sizex, sizey, render_buffer)) return SFE_FAILURE;
    const size_t npix = (size_t)sizex * sizey;
    if (sliced && vis_off + npix * 16 > c->peer.bytes) {
        set_error(fn, "the peer window holds %llu bytes; a %u x %u frame in multi-rank mode needs sfe_sim_peer_bytes_vis() = %llu",
                  (unsigned long long)c->peer.bytes, sizex, sizey, (unsigned long long)(vis_off + npix * 16));
        return SFE_FAILURE;
    }
    if (c->vis_atomic.reserve(npix * 16, fn) || c->vis_render.reserve(npix * 16, fn) || c->vis_info.reserve(sizeof(VisInfo), fn))
        return SFE_FAILURE;
    if (launch_vis(c, fn, pm4, pot, N, step_number, mat_vm, ortho_scale, autoscale, sizex, sizey, samples, pixel_value_scale,
                   tonemap, c->vis_atomic.as<int32_t>(), c->vis_render.as<float>(), c->vis_info.as<VisInfo>(), c->stream,
                   VisSlice{sliced, i0, i0 + ni, vis_off}))
        return SFE_FAILURE;
    return vis_fetch(c, fn, sizex, sizey, atomic_buffer, render_buffer, average_radius, mvp);
}

extern "C" int sfe_dev_vis_draw(const void* pm4, const float* pot, uint32_t N, uint32_t step_number, const float mat_vm[16],
                                float ortho_scale, int autoscale, uint32_t sizex, uint32_t sizey, uint32_t samples,
                                double pixel_value_scale, int tonemap, int32_t* atomic_buffer, float* render_buffer,
                                void* info, void* stream) {
    const char* fn = "sfe_dev_vis_draw";
    if (ensure_init(fn)) return SFE_FAILURE;
    if (!pm4 || !pot || !atomic_buffer || !info) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if (vis_args_ok(fn, N, mat_vm, sizex, sizey, render_buffer)) return SFE_FAILURE;
    return launch_vis(ctx(), fn, (const float4*)pm4, pot, N, step_number, mat_vm, ortho_scale, autoscale, sizex, sizey, samples,
                      pixel_value_scale, tonemap, atomic_buffer, render_buffer, (VisInfo*)info, (cudaStream_t)stream);
}
