// sfe_sph_cells.cu - cell-list (neighbour-search) evaluation of the SPH density and pressure-force sums.
//
// The reference sums every pair (sph.c:35-145 has no cutoff at all, sph.c:231 only skips r^2 > 1), but both
// kernels decay like exp(-r^2/h^2): beyond R = 4 h_max a pair contributes less than e^-16 = 1.1e-7 of a
// neighbour at r = 0, and the neglected tail of the kernel integral is 5e-7 of the density.  That is inside
// the parity gate (<= 1e-5 rel-RMS), so when R is small against the particle cloud the sums only need the
// sources in the cells that can hold a particle within R of the target's cell.  (The 1e-7 .. 5e-7 figures are relative to a
// neighbour at r = 0, i.e. they hold where the density varies slowly over R - the reference's initial conditions and the
// runs that grow out of them.  A target in a near-empty region next to a clump of N_far particles just beyond R can lose up
// to N_far e^-16 of ITS density, which no geometric test sees: the <= 1e-5 parity of this path is an RMS statement for such
// quasi-uniform inputs, checked on the test ICs.  SFE_SPH_ALL_PAIRS=1 selects the untruncated all-pairs kernels for
// anything else.)
//
//   1. bounding box + h_min/h_max (one grid-stride reduction), grid descriptor chosen ON THE DEVICE (cell size
//      s = max(R/4, cube root of 32 particles' volume), at most 2^18 cells); use_cells = 0 when the neighbourhood
//      would cover most of the box anyway - then these kernels retire at once and the all-pairs kernels of
//      sfe_sph.cu, which are enqueued behind the same flag, do the work (no host round trip either way);
//   2. stable counting sort of the particles by cell (two 9-bit LSD passes, warp-private histograms +
//      __match_any ranking, so the order inside a cell is ascending particle index and every sum is
//      bit-reproducible), cell_start[] by a scan over the per-cell counts;
//   3. 32-byte source records in cell order: {x,y,z,q0 | k,q1,own,index};
//   4. persistent CTAs draw target cells from a ticket.  For one cell the reachable sources are at most
//      (2 cpr + 1)^2 x-runs, each CONTIGUOUS in the sorted array, so they stream through shared memory as
//      plain 1-D TMA bulk copies (cp.async.bulk + mbarrier, double buffered).  A thread owns a PAIR of
//      targets in the two lanes of packed fp32x2 instructions; every lane of a warp reads the same source
//      (broadcast LDS.128).  No pair test in the inner loop: sources beyond R inside a reachable cell are
//      simply included (they only move the result towards the reference's untruncated sum).
//
// Target slices [i0, i0+ni): the sort is stable, so the owned targets of a cell are a contiguous
// sub-run of its slots.
#include "sfe_internal.cuh"
#include "sfe_scan.cuh"
#include <algorithm>
#include <stdlib.h>

namespace sfe {

constexpr uint32_t CL_CELLS_MAX = 1u << 18;
constexpr int CL_SORT_BITS = 9, CL_SORT_BINS = 1 << CL_SORT_BITS;     // 2 passes cover 2^18 cells
constexpr int CL_SORT_WARPS = 8, CL_SORT_ROWS = 16;
constexpr int CL_SORT_CHUNK = CL_SORT_WARPS * 32 * CL_SORT_ROWS;       // keys per CTA
constexpr int CL_CPR_MAX = 8;                                          // reach in cells
constexpr int CL_RUNS_MAX = (2 * CL_CPR_MAX + 1) * (2 * CL_CPR_MAX + 1);
constexpr int CL_THREADS = 64;                                         // 2 targets per thread
constexpr int CL_CHUNK = 128;                                          // sources per shared-memory stage
constexpr int CL_STAGES = 2;
constexpr uint32_t CL_SPLIT_MAX = 8;                                   // at most this many CTAs per target group (the source runs are dealt out)
constexpr uint32_t CL_SPLIT_CTAS = 16384;                              // ... and only until a pass has about this many CTAs
constexpr float CL_KAPPA = 4.0f;                                       // R = CL_KAPPA * h_max
constexpr float CL_TARGET_PER_CELL = 32.0f;                            // (default; see cl_tuning())
constexpr uint32_t CL_MIN_N = 8192;                                    // below this the all-pairs kernel is as fast
constexpr float CL_FAR = 1.0e15f;

// Cell size: s = max(R / div, edge of a cube that holds `tpc` particles on average).  Finer cells follow the sphere of
// radius R more closely (fewer pairs evaluated beyond R) at the price of more, shorter source runs.
// SFE_SPH_CELL_DIV / SFE_SPH_TPC / SFE_SPH_CULL override the defaults (read once).
struct ClTuning { float div, tpc; int cull; };
static ClTuning cl_tuning() {
    static ClTuning t = {0.f, 0.f, -1};
    if (t.cull < 0) {
        const char* e;
        t.div = (e = getenv("SFE_SPH_CELL_DIV")) ? (float)atof(e) : 4.0f;
        t.tpc = (e = getenv("SFE_SPH_TPC")) ? (float)atof(e) : CL_TARGET_PER_CELL;
        t.cull = (e = getenv("SFE_SPH_CULL")) ? atoi(e) : 1;
        if (!(t.div >= 1.0f)) t.div = 1.0f;
        if (t.div > (float)CL_CPR_MAX) t.div = (float)CL_CPR_MAX;
        if (!(t.tpc >= 1.0f)) t.tpc = 1.0f;
    }
    return t;
}

// monotone float <-> uint mapping for atomicMin/atomicMax on floats of either sign
__device__ __forceinline__ uint32_t f2ord(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

// ---------------------------------------------------------------------------------------- 1. descriptor
__global__ void cl_reset_kernel(CellGrid* g) {
    for (int k = 0; k < 3; k++) { g->bmin[k] = 0xFFFFFFFFu; g->bmax[k] = 0u; }
    g->hmin = 0xFFFFFFFFu; g->hmax = 0u;
    g->use_cells = 0u;
}

__global__ void cl_bbox_kernel(const float4* __restrict__ pm4, const float* __restrict__ rad, uint32_t N, CellGrid* g) {
    uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u}, hl = 0xFFFFFFFFu, hh = 0u;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        const float4 p = pm4[j];
        const uint32_t o[3] = {f2ord(p.x), f2ord(p.y), f2ord(p.z)};
#pragma unroll
        for (int k = 0; k < 3; k++) { lo[k] = min(lo[k], o[k]); hi[k] = max(hi[k], o[k]); }
        const uint32_t hb = f2ord(rad[j]);
        hl = min(hl, hb); hh = max(hh, hb);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        lo[k] = __reduce_min_sync(0xFFFFFFFFu, lo[k]); hi[k] = __reduce_max_sync(0xFFFFFFFFu, hi[k]);
    }
    hl = __reduce_min_sync(0xFFFFFFFFu, hl); hh = __reduce_max_sync(0xFFFFFFFFu, hh);
    if ((threadIdx.x & 31u) == 0u) {
#pragma unroll
        for (int k = 0; k < 3; k++) { atomicMin(&g->bmin[k], lo[k]); atomicMax(&g->bmax[k], hi[k]); }
        atomicMin(&g->hmin, hl); atomicMax(&g->hmax, hh);
    }
}

__global__ void cl_grid_kernel(CellGrid* g, uint32_t N, uint32_t force_all_pairs, float cell_div, float tpc) {
    const float hmax = ord2f(g->hmax), hmin = ord2f(g->hmin);
    g->uniform_h = g->hmax == g->hmin ? 1u : 0u;
    float lo[3], ext[3];
    bool finite = isfinite(hmax) && hmin > 0.0f;
    for (int k = 0; k < 3; k++) {
        lo[k] = ord2f(g->bmin[k]);
        ext[k] = ord2f(g->bmax[k]) - lo[k];
        finite = finite && isfinite(lo[k]) && isfinite(ext[k]);
    }
    if (!finite || force_all_pairs || N < CL_MIN_N) { g->use_cells = 0u; return; }
    const float R = CL_KAPPA * hmax;
    const float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
    const float tiny = fmaxf(emax * 1.0e-3f, 1.0e-30f);
    float vol = 1.0f;
    for (int k = 0; k < 3; k++) vol *= fmaxf(ext[k], tiny);
    float s = fmaxf(R / cell_div, cbrtf(vol * tpc / (float)N));
    uint32_t n[3];
    for (int it = 0; it < 200; it++) {
        uint64_t cells = 1;
        for (int k = 0; k < 3; k++) { n[k] = (uint32_t)fminf(floorf(ext[k] / s), 1.0e6f) + 1u; cells *= n[k]; }
        if (cells <= CL_CELLS_MAX) break;
        s *= 1.25f;
    }
    const uint32_t cpr = (uint32_t)ceilf(R / s);
    // work estimate: neighbourhood volume / box volume
    float frac = 1.0f;
    for (int k = 0; k < 3; k++) frac *= fminf(1.0f, (float)(2u * cpr + 1u) / (float)n[k]);
    g->origin[0] = lo[0]; g->origin[1] = lo[1]; g->origin[2] = lo[2];
    g->inv_s = 1.0f / s;
    g->rr = (R / s) * (R / s);
    g->n[0] = n[0]; g->n[1] = n[1]; g->n[2] = n[2];
    g->ncells = n[0] * n[1] * n[2];
    g->cpr = cpr;
    g->use_cells = (cpr >= 1u && cpr <= (uint32_t)CL_CPR_MAX && frac < 0.45f && (uint64_t)n[0] * n[1] * n[2] <= CL_CELLS_MAX) ? 1u : 0u;
}

// ---------------------------------------------------------------------------------------- 2. sort by cell
__global__ void cl_cellid_kernel(const float4* __restrict__ pm4, uint32_t N, const CellGrid* __restrict__ g,
                                 uint32_t* __restrict__ key, uint32_t* __restrict__ val, uint32_t* __restrict__ count) {
    if (!g->use_cells) return;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const float4 p = pm4[j];
    const float q[3] = {p.x, p.y, p.z};
    uint32_t c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float f = floorf((q[k] - g->origin[k]) * g->inv_s);
        c[k] = (uint32_t)fminf(fmaxf(f, 0.0f), (float)(g->n[k] - 1u));
    }
    const uint32_t cell = c[0] + g->n[0] * (c[1] + g->n[1] * c[2]);
    key[j] = cell; val[j] = j;
    atomicAdd(&count[cell], 1u);
}

struct CellScanOp {   // exclusive scan of count[0 .. ncells] -> start[0 .. ncells]  (entry ncells = total)
    const uint32_t* count; uint32_t* start; const CellGrid* g;
    __device__ uint32_t n() const { return g->use_cells ? g->ncells + 1u : 0u; }
    __device__ uint32_t load(uint32_t i) const { return i < g->ncells ? count[i] : 0u; }
    __device__ void store(uint32_t i, uint32_t ex) const { start[i] = ex; }
    __device__ void total(uint32_t) const {}
};

__global__ void __launch_bounds__(256) cl_sort_hist_kernel(const uint32_t* __restrict__ key, uint32_t N, int shift,
                                                           uint32_t nblocks, const CellGrid* __restrict__ g,
                                                           uint32_t* __restrict__ blockhist) {
    if (!g->use_cells) return;
    __shared__ uint32_t h[CL_SORT_BINS];
    for (int b = threadIdx.x; b < CL_SORT_BINS; b += blockDim.x) h[b] = 0u;
    __syncthreads();
    const uint32_t base = blockIdx.x * CL_SORT_CHUNK;
    for (uint32_t k = threadIdx.x; k < (uint32_t)CL_SORT_CHUNK; k += blockDim.x)
        if (base + k < N) atomicAdd(&h[(key[base + k] >> shift) & (CL_SORT_BINS - 1)], 1u);
    __syncthreads();
    for (int b = threadIdx.x; b < CL_SORT_BINS; b += blockDim.x) blockhist[(size_t)b * nblocks + blockIdx.x] = h[b];
}

struct HistScanOp {   // in-place exclusive scan of the bin-major (bin, block) count matrix
    uint32_t* m; uint32_t len; const CellGrid* g;
    __device__ uint32_t n() const { return g->use_cells ? len : 0u; }
    __device__ uint32_t load(uint32_t i) const { return m[i]; }
    __device__ void store(uint32_t i, uint32_t ex) const { m[i] = ex; }
    __device__ void total(uint32_t) const {}
};

__global__ void __launch_bounds__(CL_SORT_WARPS * 32)
cl_sort_scatter_kernel(const uint32_t* __restrict__ key, const uint32_t* __restrict__ val, uint32_t N, int shift,
                       uint32_t nblocks, const CellGrid* __restrict__ g, const uint32_t* __restrict__ blockhist,
                       uint32_t* __restrict__ key_out, uint32_t* __restrict__ val_out) {
    if (!g->use_cells) return;
    __shared__ uint32_t woff[CL_SORT_WARPS][CL_SORT_BINS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < CL_SORT_WARPS * CL_SORT_BINS; b += blockDim.x) (&woff[0][0])[b] = 0u;
    __syncthreads();
    // every warp owns CL_SORT_ROWS consecutive rows of 32 keys
    const uint32_t wbase = blockIdx.x * CL_SORT_CHUNK + warp * (32u * CL_SORT_ROWS);
    uint32_t k[CL_SORT_ROWS], v[CL_SORT_ROWS];
#pragma unroll
    for (int r = 0; r < CL_SORT_ROWS; r++) {
        const uint32_t i = wbase + r * 32u + lane;
        k[r] = i < N ? key[i] : 0xFFFFFFFFu;
        v[r] = i < N ? val[i] : 0u;
        if (i < N) atomicAdd(&woff[warp][(k[r] >> shift) & (CL_SORT_BINS - 1)], 1u);
    }
    __syncthreads();
    // counts -> offsets: global base of (bin, block) + counts of the warps before this one
    for (int b = threadIdx.x; b < CL_SORT_BINS; b += blockDim.x) {
        uint32_t run = blockhist[(size_t)b * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < CL_SORT_WARPS; w++) { const uint32_t cnt = woff[w][b]; woff[w][b] = run; run += cnt; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < CL_SORT_ROWS; r++) {
        const uint32_t i = wbase + r * 32u + lane;
        const bool valid = i < N;
        const uint32_t vm = __ballot_sync(0xFFFFFFFFu, valid);
        const uint32_t bin = (k[r] >> shift) & (CL_SORT_BINS - 1);
        uint32_t peers = 0u, dst = 0u;
        if (valid) {
            peers = __match_any_sync(vm, bin);
            dst = woff[warp][bin] + __popc(peers & ((1u << lane) - 1u));
        }
        __syncwarp();
        if (valid && (peers & ((1u << lane) - 1u)) == 0u) woff[warp][bin] += __popc(peers);
        __syncwarp();
        if (valid) { key_out[dst] = k[r]; val_out[dst] = v[r]; }
    }
}

// ---------------------------------------------------------------------------------------- 3. records
#define SFE_INV_SQRT_PI 0.5641895835477562869480794515607725858440506293289988568440857217
#define SFE_POW_PI_3_2 5.5683279968317078452848179821188357020136243902832439107536758188
#define SFE_LOG2E 1.4426950408889634073599246810018921374266459541529859341354494069
__device__ __forceinline__ float cl_coef_k(float h) { return (float)(-SFE_LOG2E / ((double)h * (double)h)); }
__device__ __forceinline__ float cl_coef_a3half(float h) { double a = SFE_INV_SQRT_PI / (double)h; return (float)(0.5 * a * a * a); }
__device__ __forceinline__ float cl_coef_d(float h) { double h2 = (double)h * (double)h; return (float)(-1.0 / (SFE_POW_PI_3_2 * h2 * h2 * (double)h)); }

// mode 0: {x,y,z,m | k, B = 0.5 a^3 m, A = 0.5 a^3, index};  mode 1: {x,y,z,c = P/rho^2 | k, D, -, index}
__global__ void cl_rec_kernel(const float4* __restrict__ pm4, const float* __restrict__ rad, const float* __restrict__ rho,
                              const float* __restrict__ prs, int mode, const uint32_t* __restrict__ sidx, uint32_t N,
                              const CellGrid* __restrict__ g, float4* __restrict__ rec) {
    if (!g->use_cells) return;
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= N) return;
    const uint32_t j = sidx[slot];
    const float4 p = pm4[j];
    const float h = rad[j];
    if (mode == 0) {
        const float a3 = cl_coef_a3half(h);
        rec[2u * slot] = p;
        rec[2u * slot + 1u] = make_float4(cl_coef_k(h), a3 * p.w, a3, __uint_as_float(j));
    } else {
        const float r2 = __fmul_rn(rho[j], rho[j]);
        const float cj = 0.f < r2 ? __fdiv_rn(prs[j], r2) : 0.f;   // sph.c:194,226
        rec[2u * slot] = make_float4(p.x, p.y, p.z, cj);
        rec[2u * slot + 1u] = make_float4(cl_coef_k(h), cl_coef_d(h), 0.f, __uint_as_float(j));
    }
}

// ---------------------------------------------------------------------------------------- 4. pair kernel
template <int MODE, bool UNIFORM_H>
__device__ __forceinline__ void cl_chunk(const float4* __restrict__ sm, uint32_t cnt, const float2 nx, const float2 ny,
                                         const float2 nz, const float2 ki, const float2 qi, const float2 ci, float cutoff2,
                                         float2& a0, float2& a1, float2& a2) {
#pragma unroll 4
    for (uint32_t j = 0; j < cnt; j++) {
        const float4 A = sm[2 * j], B = sm[2 * j + 1];
        // d' = r_j - r_i (targets are stored negated); the force sum needs r_i - r_j, undone by the caller
        const float2 dx = __fadd2_rn(nx, make_float2(A.x, A.x));
        const float2 dy = __fadd2_rn(ny, make_float2(A.y, A.y));
        const float2 dz = __fadd2_rn(nz, make_float2(A.z, A.z));
        float2 r2 = __fmul2_rn(dx, dx);
        r2 = __ffma2_rn(dy, dy, r2);
        r2 = __ffma2_rn(dz, dz, r2);
        const float2 ti = __fmul2_rn(r2, ki);
        float2 ei, ej;
        ei.x = ex2_approx(ti.x); ei.y = ex2_approx(ti.y);
        if (UNIFORM_H) ej = ei;
        else {
            const float2 tj = __fmul2_rn(r2, make_float2(B.x, B.x));
            ej.x = ex2_approx(tj.x); ej.y = ex2_approx(tj.y);
        }
        if (MODE == 0) {
            a0 = __ffma2_rn(ei, make_float2(A.w, A.w), a0);      // sum m_j e_i
            a1 = __ffma2_rn(ej, make_float2(B.y, B.y), a1);      // sum B_j e_j
        } else {
            const float2 cs = __fadd2_rn(ci, make_float2(A.w, A.w));
            float2 ds;
            if (UNIFORM_H) ds = __fmul2_rn(__fadd2_rn(qi, make_float2(B.y, B.y)), ei);
            else ds = __ffma2_rn(qi, ei, __fmul2_rn(make_float2(B.y, B.y), ej));
            float2 gg = __fmul2_rn(cs, ds);
            gg.x = r2.x > cutoff2 ? 0.f : gg.x;                   // sph.c:231
            gg.y = r2.y > cutoff2 ? 0.f : gg.y;
            a0 = __ffma2_rn(gg, dx, a0); a1 = __ffma2_rn(gg, dy, a1); a2 = __ffma2_rn(gg, dz, a2);
        }
    }
}

// Owned targets of a slice in slot order (the whole range needs no list): flag scan + compaction.
struct TargetScanOp {
    const float4* rec; uint32_t* tlist; uint32_t N, i0, ni; const CellGrid* g;
    __device__ uint32_t n() const { return g->use_cells ? N : 0u; }
    __device__ uint32_t load(uint32_t slot) const { const uint32_t j = __float_as_uint(rec[2u * slot + 1u].w); return (j >= i0 && j - i0 < ni) ? 1u : 0u; }
    __device__ void store(uint32_t slot, uint32_t ex) const { if (load(slot)) tlist[ex] = slot; }
    __device__ void total(uint32_t) const {}
};

// One CTA = one GROUP of up to 128 targets that are consecutive in cell order (thread t owns targets 2t, 2t+1 in
// the two packed lanes), so every lane does useful work whatever the cell occupancies are.  A group usually lies
// inside one x-row of cells; it is processed one row at a time: the sources of a row segment [xa, xb] are, for
// every neighbouring row (dy, dz), ONE contiguous run of the sorted records, cells [xa - amax, xb + amax].
template <int MODE>
__global__ void __launch_bounds__(CL_THREADS)
cl_pair_kernel(const CellGrid* __restrict__ g, const uint32_t* __restrict__ start, const uint32_t* __restrict__ key,
               const float4* __restrict__ rec, const uint32_t* __restrict__ tlist, uint32_t slot0, uint32_t nt,
               const float4* __restrict__ pm4, float cutoff2, float* __restrict__ out, SphRoute route, uint32_t N,
               float4* __restrict__ part, uint32_t* __restrict__ arrive, unsigned long long* __restrict__ stat, int cull) {
    if (!g->use_cells) return;
    // C3-sized problems (N = 2^18: 2048 groups of 128 targets) do not even fill the machine once, so a pass lasts as long
    // as its slowest group while the cheap groups near the surface leave their SMs idle.  Every group is therefore worked
    // on by gridDim.y CTAs (1..CL_SPLIT_MAX, chosen from the number of groups) that deal the neighbouring source runs out
    // between them; each stores its partial Kahan accumulators and the one that arrives LAST adds them up in a fixed order
    // (part 0, 1, ...) and writes the result - more, shorter CTAs and still one reproducible sum per target.
    // [B200] C3 solve_sph: 10.5 ms unsplit, 9.5 / 8.8 / 8.3 ms with 2 / 4 / 8 CTAs per group.
    const uint32_t half = blockIdx.y, nsplit = gridDim.y;
    __shared__ uint32_t s_last;
    __shared__ __align__(128) float4 tile[CL_STAGES][CL_CHUNK * 2];
    __shared__ __align__(8) uint64_t full[CL_STAGES];
    __shared__ uint32_t run_lo[CL_RUNS_MAX], run_len[CL_RUNS_MAX];
    __shared__ float run_rem[CL_RUNS_MAX];             // (R/s)^2 minus the squared cell distance in y and z of the run's row
    __shared__ float ch_rem[CL_STAGES];                // ... of the staged chunk's run
    __shared__ uint32_t s_x[2];
    unsigned long long npairs = 0ull;                   // pairs this thread evaluates (statistics)
    const uint32_t tid = threadIdx.x;
    const uint32_t nx_ = g->n[0], ny_ = g->n[1], nz_ = g->n[2];
    const int cpr = (int)g->cpr;
    const int w = 2 * cpr + 1;
    const uint32_t nruns = (uint32_t)(w * w);
    const float rr = g->rr;
    const float org_x = g->origin[0], inv_s = g->inv_s;
    const bool uniform_h = g->uniform_h != 0u;
    if (tid == 0) {
        for (int s = 0; s < CL_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t gc = 0;   // chunks consumed so far by this CTA: stage = gc % STAGES, parity = (gc / STAGES) & 1
    uint32_t s_run = 0;   // run of the chunk next_chunk() returned last

    auto next_chunk = [&](uint32_t& r, uint32_t& off, uint32_t& slot, uint32_t& cnt) -> bool {
        while (r < nruns && off >= run_len[r]) { r++; off = 0; }
        if (r >= nruns) return false;
        s_run = r;
        cnt = min((uint32_t)CL_CHUNK, run_len[r] - off);
        slot = run_lo[r] + off;
        off += cnt;
        return true;
    };

    // this CTA's targets
    const uint32_t g0 = blockIdx.x * (2u * CL_THREADS);
    const uint32_t tcount = min(2u * CL_THREADS, nt - g0);
    const uint32_t ia = 2u * tid, ib = ia + 1u;
    const bool va = ia < tcount, vb = ib < tcount;
    const uint32_t sa = va ? (tlist ? tlist[g0 + ia] : slot0 + g0 + ia) : 0u, sb = vb ? (tlist ? tlist[g0 + ib] : slot0 + g0 + ib) : 0u;
    float4 Aa = make_float4(0.f, 0.f, 0.f, 0.f), Ba = make_float4(-1.f, 0.f, 0.f, 0.f), Ab = Aa, Bb = Ba;
    uint32_t ca = 0xFFFFFFFFu, cb = 0xFFFFFFFFu;
    if (va) { Aa = rec[2u * sa]; Ba = rec[2u * sa + 1u]; ca = key[sa]; }
    if (vb) { Ab = rec[2u * sb]; Bb = rec[2u * sb + 1u]; cb = key[sb]; }
    const uint32_t rowa = va ? ca / nx_ : 0xFFFFFFFFu, rowb = vb ? cb / nx_ : 0xFFFFFFFFu;
    const uint32_t slot_first = tlist ? tlist[g0] : slot0 + g0, slot_last = tlist ? tlist[g0 + tcount - 1u] : slot0 + g0 + tcount - 1u;
    const uint32_t row_first = key[slot_first] / nx_, row_last = key[slot_last] / nx_;

    const float2 ki = make_float2(Ba.x, Bb.x);
    const float2 qi = MODE == 0 ? make_float2(Ba.z, Bb.z) : make_float2(Ba.y, Bb.y);
    const float2 ci = make_float2(Aa.w, Ab.w);
    Kahan tot[2][3];
#pragma unroll
    for (int l = 0; l < 2; l++)
#pragma unroll
        for (int q = 0; q < 3; q++) tot[l][q].s = tot[l][q].c = 0.f;

    for (uint32_t row = row_first; row <= row_last; row++) {
        // x extent of the group's targets in this row of cells
        __syncthreads();
        if (tid == 0) { s_x[0] = 0xFFFFFFFFu; s_x[1] = 0u; }
        __syncthreads();
        const bool aa = rowa == row, ab = rowb == row;
        if (aa) { atomicMin(&s_x[0], ca % nx_); atomicMax(&s_x[1], ca % nx_); }
        if (ab) { atomicMin(&s_x[0], cb % nx_); atomicMax(&s_x[1], cb % nx_); }
        __syncthreads();
        const uint32_t xa = s_x[0], xb = s_x[1];
        if (xa == 0xFFFFFFFFu) continue;          // no target of the group in this row
        const uint32_t cy = row % ny_, cz = row / ny_;
        for (uint32_t r = tid; r < nruns; r += CL_THREADS) {
            const int dy = (int)(r % (uint32_t)w) - cpr, dz = (int)(r / (uint32_t)w) - cpr;
            const int yy = (int)cy + dy, zz = (int)cz + dz;
            uint32_t rl = 0, rn = 0;
            float remv = 0.f;
            if (yy >= 0 && yy < (int)ny_ && zz >= 0 && zz < (int)nz_) {
                const float ty = (float)max(abs(dy) - 1, 0), tz = (float)max(abs(dz) - 1, 0);
                const float rem = rr - ty * ty - tz * tz;
                remv = rem;
                if (rem > 0.f) {
                    const int amax = (int)ceilf(sqrtf(rem));
                    const int x0 = max((int)xa - amax, 0), x1 = min((int)xb + amax, (int)nx_ - 1);
                    const uint32_t rbase = nx_ * ((uint32_t)yy + ny_ * (uint32_t)zz);
                    rl = start[rbase + (uint32_t)x0];
                    rn = start[rbase + (uint32_t)x1 + 1u] - rl;
                }
            }
            run_lo[r] = rl; run_len[r] = (r % nsplit) == half ? rn : 0u; run_rem[r] = remv;
        }
        __syncthreads();

        // targets of other rows sit far away in this pass: they collect exact zeros
        const float2 nx = make_float2(aa ? -Aa.x : -CL_FAR, ab ? -Ab.x : -CL_FAR);
        const float2 ny = make_float2(aa ? -Aa.y : -CL_FAR, ab ? -Ab.y : -CL_FAR);
        const float2 nz = make_float2(aa ? -Aa.z : -CL_FAR, ab ? -Ab.z : -CL_FAR);
        const bool warp_active = __any_sync(0xFFFFFFFFu, aa || ab);
        // x cells of this WARP's targets in the row: a staged chunk whose cells all lie farther than R from every one of them
        // (whole cells apart, like the run selection) is skipped by the warp
        const uint32_t wxa = __reduce_min_sync(0xFFFFFFFFu, min(aa ? ca % nx_ : 0xFFFFFFFFu, ab ? cb % nx_ : 0xFFFFFFFFu));
        const uint32_t wxb = __reduce_max_sync(0xFFFFFFFFu, max(aa ? ca % nx_ : 0u, ab ? cb % nx_ : 0u));

        uint32_t pr = 0, poff = 0, cr = 0, coff = 0, slot = 0, cnt = 0;
        if (tid == 0) {
            for (uint32_t s = 0; s < (uint32_t)CL_STAGES; s++) {
                uint32_t ps, pc;
                if (!next_chunk(pr, poff, ps, pc)) break;
                uint64_t* bar = &full[(gc + s) % CL_STAGES];
                ch_rem[(gc + s) % CL_STAGES] = run_rem[s_run];       // (before the arrive: released with it)
                mbar_expect_tx(bar, pc * 32u);
                bulk_g2s(&tile[(gc + s) % CL_STAGES][0], rec + 2u * (size_t)ps, pc * 32u, bar);
            }
        }
        while (next_chunk(cr, coff, slot, cnt)) {
            const uint32_t st = gc % CL_STAGES;
            mbar_wait(&full[st], (gc / CL_STAGES) & 1u);
            bool near = true;
            if (cull) {
                // the chunk is a piece of one x-run, sorted by cell: its x cells span those of its first and last source
                // (same expression as cl_cellid_kernel)
                const float xf = tile[st][0].x, xl = tile[st][2u * (cnt - 1u)].x;
                const uint32_t c0 = (uint32_t)fminf(fmaxf(floorf((xf - org_x) * inv_s), 0.0f), (float)(nx_ - 1u));
                const uint32_t c1 = (uint32_t)fminf(fmaxf(floorf((xl - org_x) * inv_s), 0.0f), (float)(nx_ - 1u));
                const float gap = c0 > wxb + 1u ? (float)(c0 - wxb - 1u) : (wxa > c1 + 1u ? (float)(wxa - c1 - 1u) : 0.f);
                near = gap * gap < ch_rem[st];
            }
            if (warp_active && near) {
                npairs += (unsigned long long)cnt * ((aa ? 1u : 0u) + (ab ? 1u : 0u));
                float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0;
                if (uniform_h) cl_chunk<MODE, true>(&tile[st][0], cnt, nx, ny, nz, ki, qi, ci, cutoff2, a0, a1, a2);
                else cl_chunk<MODE, false>(&tile[st][0], cnt, nx, ny, nz, ki, qi, ci, cutoff2, a0, a1, a2);
                if (MODE == 0) {
                    if (aa) { tot[0][0].add(__fmul_rn(qi.x, a0.x)); tot[0][1].add(a1.x); }
                    if (ab) { tot[1][0].add(__fmul_rn(qi.y, a0.y)); tot[1][1].add(a1.y); }
                } else {
                    if (aa) { tot[0][0].add(a0.x); tot[0][1].add(a1.x); tot[0][2].add(a2.x); }
                    if (ab) { tot[1][0].add(a0.y); tot[1][1].add(a1.y); tot[1][2].add(a2.y); }
                }
            }
            __syncthreads();   // everyone is done with stage st
            if (tid == 0) {
                uint32_t ps, pc;
                if (next_chunk(pr, poff, ps, pc)) {
                    ch_rem[st] = run_rem[s_run];
                    mbar_expect_tx(&full[st], pc * 32u);
                    bulk_g2s(&tile[st][0], rec + 2u * (size_t)ps, pc * 32u, &full[st]);
                }
            }
            gc++;
        }
    }
    if (stat) {                                             // statistics: pairs evaluated by this pass
#pragma unroll
        for (int sh = 16; sh >= 1; sh >>= 1) npairs += __shfl_xor_sync(0xFFFFFFFFu, npairs, sh);
        if ((tid & 31u) == 0u && npairs) atomicAdd(stat, npairs);
    }
    // hand-over between the gridDim.y CTAs of the group: partials to global memory, the last arrival combines
    float4* mine = part + ((size_t)(blockIdx.x * nsplit + half) * 3u) * CL_THREADS;
#pragma unroll
    for (int q = 0; q < 3; q++) mine[q * CL_THREADS + tid] = make_float4(tot[0][q].s, tot[0][q].c, tot[1][q].s, tot[1][q].c);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const uint32_t seen = atomicAdd(&arrive[blockIdx.x], 1u);
        s_last = seen == nsplit - 1u ? 1u : 0u;
        if (s_last) arrive[blockIdx.x] = 0u;          // ready for the next pass
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double dsum[2] = {0.0, 0.0};                      // MODE 0: rho in double, half by half
    float fsum[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    for (uint32_t h = 0; h < nsplit; h++) {
        const float4* theirs = part + ((size_t)(blockIdx.x * nsplit + h) * 3u) * CL_THREADS;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const float4 v = __ldcg(&theirs[q * CL_THREADS + tid]);
            if (MODE == 0) {
                if (q < 2) { dsum[0] += (double)(v.x - v.y); dsum[1] += (double)(v.z - v.w); }
            } else {
                fsum[0][q] = __fadd_rn(fsum[0][q], v.x - v.y); fsum[1][q] = __fadd_rn(fsum[1][q], v.z - v.w);
            }
        }
    }
#pragma unroll
    for (int l = 0; l < 2; l++) {
        if (!(l == 0 ? va : vb)) continue;
        const uint32_t i = __float_as_uint(l == 0 ? Ba.w : Bb.w);
        if (MODE == 0) {
            const float v = (float)dsum[l];
            if (route.world) {   // every rank needs rho of every source for the force pass: store it everywhere
                for (uint32_t r = 0; r < route.world; r++) reinterpret_cast<float*>(route.base[r] + route.off)[i] = v;
            } else out[i] = v;
        } else {
            if (route.world) {   // the pressure force goes into the acc array of the rank that owns particle i
                uint32_t r = (uint32_t)(((uint64_t)i * route.world) / N);
                while (i < (uint32_t)(((uint64_t)N * r) / route.world)) r--;
                while (i >= (uint32_t)(((uint64_t)N * (r + 1u)) / route.world)) r++;
                out = reinterpret_cast<float*>(route.base[r] + route.off);
            }
            // sums were taken with d' = -(r_i - r_j): acc += -m_i * S  (sph.c:274-276)  =>  acc += m_i * S'
            const float m = pm4[i].w;
            out[3u * i + 0u] = __fadd_rn(out[3u * i + 0u], __fmul_rn(m, fsum[l][0]));
            out[3u * i + 1u] = __fadd_rn(out[3u * i + 1u], __fmul_rn(m, fsum[l][1]));
            out[3u * i + 2u] = __fadd_rn(out[3u * i + 2u], __fmul_rn(m, fsum[l][2]));
        }
    }
}

// ---------------------------------------------------------------------------------------- launcher
int cells_prepare(Ctx* c, const float4* pm4, const float* rad, uint32_t N, bool force_all_pairs, cudaStream_t s, const char* fn) {
    SphCells& k = c->sph_cells;
    if (k.desc.reserve(sizeof(CellGrid), fn)) return SFE_FAILURE;
    CellGrid* g = k.desc.as<CellGrid>();
    cl_reset_kernel<<<1, 1, 0, s>>>(g);
    if (check_launch(fn)) return SFE_FAILURE;
    const uint32_t rb = std::min<uint32_t>((N + 255) / 256, 4u * (uint32_t)c->sm_count);
    cl_bbox_kernel<<<rb, 256, 0, s>>>(pm4, rad, N, g);
    if (check_launch(fn)) return SFE_FAILURE;
    cl_grid_kernel<<<1, 1, 0, s>>>(g, N, (force_all_pairs || N < CL_MIN_N) ? 1u : 0u, cl_tuning().div, cl_tuning().tpc);
    if (check_launch(fn)) return SFE_FAILURE;
    k.maybe = !(force_all_pairs || N < CL_MIN_N);
    if (!k.maybe) return SFE_SUCCESS;

    const uint32_t nblocks = (N + CL_SORT_CHUNK - 1) / CL_SORT_CHUNK;
    if (k.count.reserve((size_t)(CL_CELLS_MAX + 1) * 4, fn) || k.start.reserve((size_t)(CL_CELLS_MAX + 2) * 4, fn) ||
        k.key[0].reserve((size_t)N * 4, fn) || k.key[1].reserve((size_t)N * 4, fn) ||
        k.val[0].reserve((size_t)N * 4, fn) || k.val[1].reserve((size_t)N * 4, fn) ||
        k.hist.reserve((size_t)CL_SORT_BINS * nblocks * 4, fn) || k.rec.reserve((size_t)N * 32 + 64, fn))
        return SFE_FAILURE;
    SFE_CUDA(fn, cudaMemsetAsync(k.count.p, 0, (size_t)(CL_CELLS_MAX + 1) * 4, s));
    cl_cellid_kernel<<<(N + 255) / 256, 256, 0, s>>>(pm4, N, g, k.key[0].as<uint32_t>(), k.val[0].as<uint32_t>(), k.count.as<uint32_t>());
    if (check_launch(fn)) return SFE_FAILURE;
    if (run_scan(c, CellScanOp{k.count.as<uint32_t>(), k.start.as<uint32_t>(), g}, CL_CELLS_MAX + 1, s, fn)) return SFE_FAILURE;
    for (int pass = 0; pass < 2; pass++) {
        const int shift = pass * CL_SORT_BITS;
        const uint32_t *ki = k.key[pass].as<uint32_t>(), *vi = k.val[pass].as<uint32_t>();
        uint32_t *ko = k.key[pass ^ 1].as<uint32_t>(), *vo = k.val[pass ^ 1].as<uint32_t>();
        cl_sort_hist_kernel<<<nblocks, 256, 0, s>>>(ki, N, shift, nblocks, g, k.hist.as<uint32_t>());
        if (check_launch(fn)) return SFE_FAILURE;
        if (run_scan(c, HistScanOp{k.hist.as<uint32_t>(), CL_SORT_BINS * nblocks, g}, CL_SORT_BINS * nblocks, s, fn)) return SFE_FAILURE;
        cl_sort_scatter_kernel<<<nblocks, CL_SORT_WARPS * 32, 0, s>>>(ki, vi, N, shift, nblocks, g, k.hist.as<uint32_t>(), ko, vo);
        if (check_launch(fn)) return SFE_FAILURE;
    }
    // after two passes the sorted order is back in key[0] / val[0]
    return SFE_SUCCESS;
}

int cells_run(Ctx* c, int mode, const float4* pm4, const float* rad, const float* rho, const float* prs, uint32_t N,
              uint32_t i0, uint32_t ni, float cutoff2, float* out, cudaStream_t s, const char* fn, const SphRoute* route) {
    SphCells& k = c->sph_cells;
    if (!k.maybe) return SFE_SUCCESS;
    const CellGrid* g = k.desc.as<CellGrid>();
    cl_rec_kernel<<<(N + 255) / 256, 256, 0, s>>>(pm4, rad, rho, prs, mode, k.val[0].as<uint32_t>(), N, g, k.rec.as<float4>());
    if (check_launch(fn)) return SFE_FAILURE;
    const uint32_t* tlist = nullptr;
    SphRoute rt; rt.world = 0u; rt.rank = 0u; rt.off = 0u;
    uint32_t slot0 = 0u;
    if (route) {                   // multi-rank: a dense range of sorted slots, results routed to the owners
        rt = *route;
        slot0 = (uint32_t)(((uint64_t)N * rt.rank) / rt.world);
        ni = (uint32_t)(((uint64_t)N * (rt.rank + 1u)) / rt.world) - slot0;
        if (ni == 0u) return SFE_SUCCESS;
    } else if (!(i0 == 0 && ni == N)) {   // owned targets of the slice, in cell order
        if (k.tlist.reserve((size_t)ni * 4, fn)) return SFE_FAILURE;
        if (run_scan(c, TargetScanOp{k.rec.as<float4>(), k.tlist.as<uint32_t>(), N, i0, ni, g}, N, s, fn)) return SFE_FAILURE;
        tlist = k.tlist.as<uint32_t>();
    }
    const uint32_t groups = (ni + 2u * CL_THREADS - 1u) / (2u * CL_THREADS);
    uint32_t nsplit = groups ? CL_SPLIT_CTAS / groups : 1u;
    nsplit = nsplit < 1u ? 1u : (nsplit > CL_SPLIT_MAX ? CL_SPLIT_MAX : nsplit);
    if (k.part.reserve((size_t)groups * nsplit * 3u * CL_THREADS * sizeof(float4), fn)) return SFE_FAILURE;
    if ((size_t)groups * 4 > k.arrive.cap) {          // arrival counters: zero once, every pass leaves them at zero
        if (k.arrive.reserve((size_t)groups * 4, fn)) return SFE_FAILURE;
        SFE_CUDA(fn, cudaMemsetAsync(k.arrive.p, 0, k.arrive.cap, s));
    }
    if (!k.stat.p) {
        if (k.stat.reserve(8, fn)) return SFE_FAILURE;
        SFE_CUDA(fn, cudaMemsetAsync(k.stat.p, 0, 8, s));
    }
    const dim3 grid(groups, nsplit);
    if (mode == 0)
        cl_pair_kernel<0><<<grid, CL_THREADS, 0, s>>>(g, k.start.as<uint32_t>(), k.key[0].as<uint32_t>(), k.rec.as<float4>(), tlist, slot0, ni, pm4, cutoff2, out, rt, N,
                                                      k.part.as<float4>(), k.arrive.as<uint32_t>(), k.stat.as<unsigned long long>(), cl_tuning().cull);
    else
        cl_pair_kernel<1><<<grid, CL_THREADS, 0, s>>>(g, k.start.as<uint32_t>(), k.key[0].as<uint32_t>(), k.rec.as<float4>(), tlist, slot0, ni, pm4, cutoff2, out, rt, N,
                                                      k.part.as<float4>(), k.arrive.as<uint32_t>(), k.stat.as<unsigned long long>(), cl_tuning().cull);
    return check_launch(fn);
}

// statistics hook: pairs evaluated by the cell-list passes since the last call (read and reset)
int cells_stats(Ctx* c, unsigned long long* pairs, cudaStream_t s, const char* fn) {
    SphCells& k = c->sph_cells;
    *pairs = 0ull;
    if (!k.stat.p) return SFE_SUCCESS;
    SFE_CUDA(fn, cudaMemcpyAsync(pairs, k.stat.p, 8, cudaMemcpyDeviceToHost, s));
    SFE_CUDA(fn, cudaMemsetAsync(k.stat.p, 0, 8, s));
    SFE_CUDA(fn, cudaStreamSynchronize(s));
    return SFE_SUCCESS;
}

}  // namespace sfe
