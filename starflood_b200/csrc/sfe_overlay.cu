// sfe_overlay.cu - the SOLVER_VIS debug overlay of solve_gravity_part_tree().
//
// With SOLVER_VIS defined (the reference's default, config.h:32) the tree walk of ONE particle draws into the
// visualisation's render buffer: the outline of every internal cell it looks at, coloured by depth
// (gravity.c:380-445), and a segment from the particle to every monopole it evaluates (gravity.c:550-574), rasterised
// with `+=` by an error-term line walk (gravity.c:16-111).  The picture is a pure function of the tree and that
// particle, and it is reproduced here bit for bit on the device tree:
//
//   overlay_walk_kernel     one thread replays the particle's depth-first visit order on the engine's compact records
//                           (cell bounds, {COM, mass}, slot ranges), projects every segment's ends through matMVP and
//                           appends {x0, y0, x1, y1, rgba} to a list - the drawing order of the reference;
//   overlay_raster_kernel   the list is drawn IN ORDER (float `+=` does not commute with itself across different
//                           colours), one segment at a time, but all pixels of a segment at once: pixel k of the
//                           error-term walk has the closed form  major = a0 + s*k,  minor = b0 + s'*floor((2*k*dmin +
//                           dmaj) / (2*dmaj)), and a segment never visits a pixel twice, so the adds of one segment are
//                           independent.
//
// Only the rows of the render buffer that the segments can touch travel over PCIe.
#include "sfe_internal.cuh"
#include "sfe_tree.cuh"

namespace sfe {

struct __align__(16) OvLine { int x0, y0, x1, y1; float rgba[4]; };
struct OvInfo { uint32_t count, stored; int ymin, ymax; };
struct Mvp { float m[16]; };

__device__ __forceinline__ void ov_project(const Mvp& M, const float q[3], uint32_t sizex, uint32_t sizey, int* px, int* py) {
    float h[2];
#pragma unroll
    for (int r = 0; r < 2; r++)   // gravity.c:27-33: products rounded one by one, summed left to right, w = 1
        h[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M.m[4 * r + 0], q[0]), __fmul_rn(M.m[4 * r + 1], q[1])),
                                   __fmul_rn(M.m[4 * r + 2], q[2])), __fmul_rn(M.m[4 * r + 3], 1.0f));
    const float hy = __fmul_rn(0.5f, (float)sizey);
    // gravity.c:61-65: (int)(-0.5 + 0.5*sizey*p + 0.5*size)
    *px = __float2int_rz(__fadd_rn(__fadd_rn(-0.5f, __fmul_rn(hy, h[0])), __fmul_rn(0.5f, (float)sizex)));
    *py = __float2int_rz(__fadd_rn(__fadd_rn(-0.5f, __fmul_rn(hy, h[1])), __fmul_rn(0.5f, (float)sizey)));
}

struct OvEmit {
    const Mvp& M; uint32_t sizex, sizey; OvLine* lines; uint32_t cap; uint32_t count; int ymin, ymax;
    __device__ void segment(const float a[3], const float b[3], const float rgba[4]) {
        OvLine L;
        ov_project(M, a, sizex, sizey, &L.x0, &L.y0);
        ov_project(M, b, sizex, sizey, &L.x1, &L.y1);
#pragma unroll
        for (int c = 0; c < 4; c++) L.rgba[c] = rgba[c];
        if (count < cap) lines[count] = L;
        count++;
        ymin = min(ymin, min(L.y0, L.y1)); ymax = max(ymax, max(L.y0, L.y1));
    }
};

__global__ void overlay_walk_kernel(TreeDev t, const float4* __restrict__ pm4, uint32_t idx, float theta, Mvp M,
                                    uint32_t sizex, uint32_t sizey, OvLine* __restrict__ lines, uint32_t cap,
                                    OvInfo* __restrict__ info) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    OvEmit out{M, sizex, sizey, lines, cap, 0u, 0x7FFFFFFF, -0x7FFFFFFF};
    const float4 me = pm4[idx];
    const float ri[3] = {me.x, me.y, me.z};
    const float mi = me.w;
    // depth-first, children in octant order; `child` = -1 until the node has been looked at (gravity.c:304-460)
    uint32_t node[SFE_MAX_LEVELS + 3];
    int child[SFE_MAX_LEVELS + 3];
    int sp = 0;
    node[0] = 0u; child[0] = -1;
    const uint32_t count = t.meta[M_COUNT];
    while (sp >= 0) {
        const uint32_t n = node[sp];
        if (child[sp] >= 0) {                                   // coming back from a child
            if (child[sp] >= 7) { sp--; continue; }
            child[sp]++;
            node[sp + 1] = t.child_base[n] + (uint32_t)child[sp]; child[sp + 1] = -1; sp++;
            continue;
        }
        const uint2 sg = t.seg[n];
        const uint32_t bodies = n < count ? sg.y - sg.x : 0u;
        if (bodies == 0u) { sp--; continue; }                   // gravity.c:324-328
        const WalkRec2 cell = t.rec2[n];
        if (bodies > 8u) {
            float off[3];
#pragma unroll
            for (int k = 0; k < 3; k++) off[k] = __fsub_rn(__fmul_rn(0.5f, __fadd_rn(cell.lo[k], cell.hi[k])), ri[k]);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(off[0], off[0]), __fmul_rn(off[1], off[1])), __fmul_rn(off[2], off[2]));
            const float w = __fmul_rn((float)(1.0 / 3.0), __fadd_rn(__fadd_rn(__fsub_rn(cell.hi[0], cell.lo[0]), __fsub_rn(cell.hi[1], cell.lo[1])),
                                                                    __fsub_rn(cell.hi[2], cell.lo[2])));
            const float ratio = 0.0f < d2 ? __fdiv_rn(__fmul_rn(w, w), d2) : 0.0f;          // gravity.c:377
            // outline, colour by depth (gravity.c:382-387): 4 edges along each axis, the lower of the two other axes fastest
            const float f = __fdiv_rn(1.0f, (float)(sp + 1));
            const float rgba[4] = {__fmul_rn(f, sp % 3 == 0 ? 0.8f : 0.1f), __fmul_rn(f, sp % 3 == 1 ? 0.8f : 0.1f),
                                   __fmul_rn(f, sp % 3 == 2 ? 0.8f : 0.1f), 1.0f};
            for (int ax = 0; ax < 3; ax++) {
                const int u = ax == 0 ? 1 : 0, v = ax == 2 ? 1 : 2;
                for (int q = 0; q < 4; q++) {
                    float p0[3], p1[3];
                    p0[u] = p1[u] = (q & 1) ? cell.hi[u] : cell.lo[u];
                    p0[v] = p1[v] = (q & 2) ? cell.hi[v] : cell.lo[v];
                    p0[ax] = cell.lo[ax]; p1[ax] = cell.hi[ax];
                    out.segment(p0, p1, rgba);
                }
            }
            if (theta <= ratio) {                               // gravity.c:448-453
                child[sp] = 0;
                node[sp + 1] = t.child_base[n]; child[sp + 1] = -1; sp++;
                continue;
            }
        }
        // evaluated as a monopole: the segment goes to the (self-corrected) centre of mass (gravity.c:463-490, :571)
        const float4 cm = t.rec1[n];
        float rj[3] = {cm.x, cm.y, cm.z};
        float mj = cm.w;
        if (cell.lo[0] <= ri[0] && ri[0] <= cell.hi[0] && cell.lo[1] <= ri[1] && ri[1] <= cell.hi[1] &&
            cell.lo[2] <= ri[2] && ri[2] <= cell.hi[2] && 1u < bodies) {
#pragma unroll
            for (int k = 0; k < 3; k++) rj[k] = __fsub_rn(__fmul_rn(rj[k], mj), __fmul_rn(mi, ri[k]));
            mj = __fsub_rn(mj, mi);
            const float inv = 0.0f < mj ? __fdiv_rn(1.0f, mj) : 1.0f;
#pragma unroll
            for (int k = 0; k < 3; k++) rj[k] = __fmul_rn(rj[k], inv);
        }
        const float magenta[4] = {0.8f, 0.1f, 0.8f, 1.0f};
        out.segment(ri, rj, magenta);
        sp--;
    }
    info->count = out.count;
    info->stored = out.count < cap ? out.count : cap;
    info->ymin = out.ymin; info->ymax = out.ymax;
}

// rows [row0, row0 + nrows) of the render buffer are resident at `rb`
__global__ void __launch_bounds__(1024) overlay_raster_kernel(const OvLine* __restrict__ lines, uint32_t count,
                                                              float4* __restrict__ rb, int row0, int nrows, int sizex, int sizey) {
    for (uint32_t l = 0; l < count; l++) {
        const OvLine L = lines[l];
        const long long dx = llabs((long long)L.x1 - L.x0), dy = llabs((long long)L.y1 - L.y0);
        const int sx = L.x0 < L.x1 ? 1 : -1, sy = L.y0 < L.y1 ? 1 : -1;
        const bool xmajor = dx >= dy;
        const long long dmaj = xmajor ? dx : dy, dmin = xmajor ? dy : dx;
        // k range whose major coordinate is on screen: 0 < a0 + s*k < size
        const long long a0 = xmajor ? L.x0 : L.y0, size = xmajor ? sizex : sizey;
        const int s = xmajor ? sx : sy;
        long long k0 = s > 0 ? 1 - a0 : a0 - size + 1, k1 = s > 0 ? size - 1 - a0 : a0 - 1;
        if (k0 < 0) k0 = 0;
        if (k1 > dmaj) k1 = dmaj;
        for (long long k = k0 + threadIdx.x; k <= k1; k += blockDim.x) {
            const long long j = dmaj ? (2 * dmin * k + dmaj) / (2 * dmaj) : 0;
            const long long x = xmajor ? L.x0 + (long long)sx * k : L.x0 + (long long)sx * j;
            const long long y = xmajor ? L.y0 + (long long)sy * j : L.y0 + (long long)sy * k;
            if (0 < x && x < sizex && 0 < y && y < sizey && y >= row0 && y < row0 + nrows) {
                float4* p = rb + (size_t)(y - row0) * sizex + x;
                float4 v = *p;
                v.x = __fadd_rn(v.x, L.rgba[0]); v.y = __fadd_rn(v.y, L.rgba[1]);
                v.z = __fadd_rn(v.z, L.rgba[2]); v.w = __fadd_rn(v.w, L.rgba[3]);
                *p = v;
            }
        }
        __syncthreads();
    }
}

}  // namespace sfe

using namespace sfe;

extern "C" int sfe_bh_overlay(uint32_t idx, float theta, const float mvp[16], float* render_buffer, uint32_t sizex, uint32_t sizey) {
    const char* fn = "sfe_bh_overlay";
    if (ensure_init(fn)) return SFE_FAILURE;
    Ctx* c = ctx();
    Tree& t = c->tree;
    if (!mvp || !render_buffer) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if (!t.valid || !t.pm4_of_build) { set_error(fn, "no valid tree: call sfe_tree_build() first"); return SFE_FAILURE; }
    if (sizex == 0 || sizey == 0 || sizex > 65536u || sizey > 65536u) { set_error(fn, "render buffer of %u x %u pixels", sizex, sizey); return SFE_FAILURE; }
    if (idx >= t.N) return SFE_SUCCESS;                          // `SOLVER_VIS == idx` is never true (gravity.c:381)
    cudaStream_t s = c->stream;
    const float4* pm4 = (const float4*)t.pm4_of_build;
    // the root's moments are deferred (sfe_tree.cu); fold them if this particle takes the root as a monopole
    if (tree_root_moments(c, pm4, idx, 1, theta, false, s)) return SFE_FAILURE;
    if (tree_status(c, s)) return SFE_FAILURE;
    Mvp M;
    for (int k = 0; k < 16; k++) M.m[k] = mvp[k];
    if (!c->ov_info) {
        SFE_CUDA(fn, cudaHostAlloc(&c->ov_info, 64, cudaHostAllocMapped));
        SFE_CUDA(fn, cudaHostGetDevicePointer(&c->ov_info_dev, c->ov_info, 0));
    }
    OvInfo* info = (OvInfo*)c->ov_info;
    size_t cap = c->ov_lines.cap / sizeof(OvLine);
    if (cap < 4096) { if (c->ov_lines.reserve(16384 * sizeof(OvLine), fn)) return SFE_FAILURE; cap = c->ov_lines.cap / sizeof(OvLine); }
    for (int attempt = 0; attempt < 2; attempt++) {
        overlay_walk_kernel<<<1, 32, 0, s>>>(tree_dev(c), pm4, idx, theta, M, sizex, sizey, c->ov_lines.as<OvLine>(), (uint32_t)cap,
                                            (OvInfo*)c->ov_info_dev);
        if (check_launch(fn)) return SFE_FAILURE;
        SFE_CUDA(fn, cudaStreamSynchronize(s));
        if (info->count <= cap) break;
        if (attempt == 1) { set_error(fn, "segment list overflow (%u segments)", info->count); return SFE_FAILURE; }
        if (c->ov_lines.reserve((size_t)info->count * sizeof(OvLine), fn)) return SFE_FAILURE;
        cap = c->ov_lines.cap / sizeof(OvLine);
    }
    if (info->count == 0) return SFE_SUCCESS;
    int r0 = info->ymin < 1 ? 1 : info->ymin, r1 = info->ymax > (int)sizey - 1 ? (int)sizey - 1 : info->ymax;
    if (r0 > r1) return SFE_SUCCESS;                             // everything off screen
    const int nrows = r1 - r0 + 1;
    const size_t row_bytes = (size_t)sizex * 16u, bytes = row_bytes * nrows;
    if (c->ov_rows.reserve(bytes, fn)) return SFE_FAILURE;
    float* host_rows = render_buffer + (size_t)r0 * sizex * 4u;
    if (h2d(fn, c->ov_rows.p, host_rows, bytes, s)) return SFE_FAILURE;
    overlay_raster_kernel<<<1, 1024, 0, s>>>(c->ov_lines.as<OvLine>(), info->count, c->ov_rows.as<float4>(), r0, nrows, (int)sizex, (int)sizey);
    if (check_launch(fn)) return SFE_FAILURE;
    if (d2h(fn, host_rows, c->ov_rows.p, bytes, s)) return SFE_FAILURE;
    SFE_CUDA(fn, cudaStreamSynchronize(s));
    return SFE_SUCCESS;
}
