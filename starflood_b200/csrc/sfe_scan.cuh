// sfe_scan.cuh - three-launch exclusive scan over device-resident element counts (used by the tree build's
// compaction / child allocation and by the walk's slice selection).
#pragma once
#include "sfe_internal.cuh"

namespace sfe {

// ============================================================================ generic exclusive scan
// Three launches: per-tile sums, scan of the tile sums (one CTA), per-tile apply.  `Op` supplies
//   uint32_t load(uint32_t i)                value of element i (i < n())
//   void     store(uint32_t i, uint32_t ex)  called with the exclusive prefix of element i
//   uint32_t n()                             element count (may live in device memory)
//   void     total(uint32_t t)               called once with the grand total
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the CTA; returns the exclusive prefix, *total = CTA sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total, uint32_t* sm /*>= 33 words*/) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = warp_incl_scan(v);
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nw ? sm[lane] : 0;
        uint32_t wi = warp_incl_scan(w);
        sm[lane] = wi - w;
        if (lane == 31) sm[32] = wi;
    }
    __syncthreads();
    uint32_t ex = inc - v + sm[warp];
    *total = sm[32];
    __syncthreads();
    return ex;
}

template <class Op>
__global__ void __launch_bounds__(SC_THREADS) scan_reduce_kernel(Op op, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t sm[33];
    const uint32_t n = op.n();
    const uint32_t ntiles = (n + SC_TILE - 1) / SC_TILE;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < SC_ITEMS; k++) {
            uint32_t i = tile * SC_TILE + k * SC_THREADS + threadIdx.x;
            if (i < n) s += op.load(i);
        }
        uint32_t tot;
        block_excl_scan(s, &tot, sm);
        if (threadIdx.x == 0) tile_sums[tile] = tot;
    }
}

template <class Op>
__global__ void __launch_bounds__(1024) scan_tiles_kernel(Op op, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t sm[33];
    __shared__ uint32_t carry;
    const uint32_t n = op.n();
    const uint32_t ntiles = (n + SC_TILE - 1) / SC_TILE;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sums[i] : 0, tot;
        uint32_t ex = block_excl_scan(v, &tot, sm);
        if (i < ntiles) tile_sums[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) op.total(carry);
}

template <class Op>
__global__ void __launch_bounds__(SC_THREADS) scan_apply_kernel(Op op, const uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t sm[33];
    const uint32_t n = op.n();
    const uint32_t ntiles = (n + SC_TILE - 1) / SC_TILE;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // thread-contiguous items so that the scan order is the element order
        uint32_t v[SC_ITEMS], s = 0;
        const uint32_t first = tile * SC_TILE + threadIdx.x * SC_ITEMS;
#pragma unroll
        for (int k = 0; k < SC_ITEMS; k++) { v[k] = (first + k < n) ? op.load(first + k) : 0; s += v[k]; }
        uint32_t tot;
        uint32_t ex = block_excl_scan(s, &tot, sm) + tile_sums[tile];
#pragma unroll
        for (int k = 0; k < SC_ITEMS; k++) { if (first + k < n) op.store(first + k, ex); ex += v[k]; }
    }
}

template <class Op>
static int run_scan(Ctx* c, Op op, uint32_t n_upper, cudaStream_t s, const char* fn) {
    Tree& t = c->tree;
    const uint32_t ntiles_upper = (n_upper + SC_TILE - 1) / SC_TILE;
    if (t.scan_tmp.reserve((size_t)(ntiles_upper + 1) * 4, fn)) return SFE_FAILURE;
    uint32_t grid = ntiles_upper < (uint32_t)c->sm_count * 8 ? ntiles_upper : (uint32_t)c->sm_count * 8;
    if (grid == 0) grid = 1;
    scan_reduce_kernel<<<grid, SC_THREADS, 0, s>>>(op, t.scan_tmp.as<uint32_t>());
    if (check_launch(fn)) return SFE_FAILURE;
    scan_tiles_kernel<<<1, 1024, 0, s>>>(op, t.scan_tmp.as<uint32_t>());
    if (check_launch(fn)) return SFE_FAILURE;
    scan_apply_kernel<<<grid, SC_THREADS, 0, s>>>(op, t.scan_tmp.as<uint32_t>());
    return check_launch(fn);
}


}  // namespace sfe
