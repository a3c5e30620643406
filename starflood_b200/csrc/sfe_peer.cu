// sfe_peer.cu - the exchange step of the multi-GPU partition over NVLink peer memory.
//
// One process per GPU.  Every rank cudaMalloc's a WINDOW (payload + one control page), publishes it with a CUDA IPC
// handle and maps the windows of all other ranks (cudaIpcOpenMemHandle -> peer access over NVLink / NVSwitch).  An
// exchange is then two launches on the producer's stream and nothing on the host:
//
//   data      a kernel stores the rank's slice into the same place of EVERY rank's window (plain 16-byte stores to
//             peer pointers; the drift kernel of sfe_sim.cu does this itself, so drift + all-gather is one kernel);
//   barrier   one warp: __threadfence_system(), lane r releases flag[rank] = epoch into rank r's control page, then
//             acquires its own flag[r] until every peer's epoch has arrived.
//
// The kernel boundary orders the data stores before the flag stores; release/acquire at system scope makes them
// visible to the consumer kernels that follow the barrier in stream order on the receiving GPU.  Flags are monotone
// epochs, so a rank that runs ahead cannot be confused with a late one; it can be at most ONE exchange ahead of its
// slowest peer (it blocks in the next barrier), which is why ping-ponging two regions is enough (sfengine.h).
// A wait gives up after SFE_PEER_TIMEOUT_MS and reports through mapped host memory: a dead peer must not hang the GPU.
#include "sfe_internal.cuh"
#include <stdlib.h>

namespace sfe {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// lane r: tell rank r that this rank has finished exchange `epoch`, then wait until rank r has told us the same
__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerPtrs pp, uint64_t ctrl_off, uint32_t epoch, uint64_t timeout_ns,
                                                          uint32_t* __restrict__ err) {
    const uint32_t r = threadIdx.x;
    __threadfence_system();
    if (r < pp.world) {
        PeerCtrl* theirs = reinterpret_cast<PeerCtrl*>(pp.base[r] + ctrl_off);
        st_release_sys(&theirs->flag[pp.rank], epoch);
        const PeerCtrl* mine = reinterpret_cast<const PeerCtrl*>(pp.base[pp.rank] + ctrl_off);
        const uint64_t t0 = globaltimer_ns();
        uint32_t spins = 0;
        while ((int32_t)(ld_acquire_sys(&mine->flag[r]) - epoch) < 0) {
            if ((++spins & 0x3FFu) == 0u && globaltimer_ns() - t0 > timeout_ns) { *err = 1u + r; __threadfence_system(); break; }
            __nanosleep(64);
        }
    }
    __syncwarp();
    __threadfence_system();
}

// [off, off+bytes) of the local window -> the same range of every peer's window: 16-byte stores for the aligned
// body, single bytes for the (at most 15-byte) head and tail
__global__ void __launch_bounds__(256) peer_put_kernel(PeerPtrs pp, uint64_t off, uint64_t bytes) {
    const char* src = pp.base[pp.rank] + off;
    uint64_t head = (16u - ((uintptr_t)src & 15u)) & 15u;
    if (head > bytes) head = bytes;
    const uint64_t n16 = (bytes - head) / 16u, tail0 = head + n16 * 16u;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = tid; k < n16; k += nth) {
        const uint4 v = reinterpret_cast<const uint4*>(src + head)[k];
        for (uint32_t r = 0; r < pp.world; r++)
            if (r != pp.rank) reinterpret_cast<uint4*>(pp.base[r] + off + head)[k] = v;
    }
    const uint64_t edge = head + (bytes - tail0);
    for (uint64_t e = tid; e < edge; e += nth) {
        const uint64_t b = e < head ? e : tail0 + (e - head);
        for (uint32_t r = 0; r < pp.world; r++)
            if (r != pp.rank) pp.base[r][off + b] = src[b];
    }
}

PeerPtrs peer_ptrs(const Ctx* c) {
    PeerPtrs pp;
    for (uint32_t r = 0; r < SFE_PEER_MAX; r++) pp.base[r] = (char*)c->peer.base[r];
    pp.world = c->peer.world; pp.rank = c->peer.rank;
    return pp;
}

static uint64_t peer_timeout_ns() {
    const char* e = getenv("SFE_PEER_TIMEOUT_MS");
    const double ms = e ? atof(e) : 10000.0;
    return (uint64_t)((ms > 1.0 ? ms : 1.0) * 1.0e6);
}

int launch_peer_barrier(Ctx* c, cudaStream_t s, const char* fn) {
    PeerState& p = c->peer;
    if (!p.open) { set_error(fn, "the peer window is not open: call sfe_peer_create() and sfe_peer_open() first"); return SFE_FAILURE; }
    p.epoch++;
    peer_barrier_kernel<<<1, 32, 0, s>>>(peer_ptrs(c), p.bytes, p.epoch, peer_timeout_ns(), p.err_dev);
    return check_launch(fn);
}

int launch_peer_put(Ctx* c, uint64_t offset, uint64_t bytes, cudaStream_t s, const char* fn) {
    PeerState& p = c->peer;
    if (!p.open) { set_error(fn, "the peer window is not open: call sfe_peer_create() and sfe_peer_open() first"); return SFE_FAILURE; }
    if (offset > p.bytes || bytes > p.bytes - offset) {
        set_error(fn, "range [%llu, +%llu) outside the %llu-byte window", (unsigned long long)offset, (unsigned long long)bytes, (unsigned long long)p.bytes);
        return SFE_FAILURE;
    }
    if (bytes && p.world > 1) {
        uint64_t g = (bytes / 16u + 255u) / 256u;
        const uint64_t cap = (uint64_t)c->sm_count * 4u;
        g = g < 1 ? 1 : (g < cap ? g : cap);
        peer_put_kernel<<<(uint32_t)g, 256, 0, s>>>(peer_ptrs(c), offset, bytes);
        if (check_launch(fn)) return SFE_FAILURE;
    }
    return launch_peer_barrier(c, s, fn);
}

int peer_status(Ctx* c, const char* fn) {
    PeerState& p = c->peer;
    if (p.err_host && *(volatile uint32_t*)p.err_host) {
        set_error(fn, "peer exchange timed out waiting for rank %u (epoch %u)", *(volatile uint32_t*)p.err_host - 1u, p.epoch);
        return SFE_FAILURE;
    }
    return SFE_SUCCESS;
}

static void peer_detach(PeerState& p) {
    if (p.open)
        for (uint32_t r = 0; r < p.world; r++)
            if (r != p.rank && p.base[r]) { cudaIpcCloseMemHandle(p.base[r]); p.base[r] = nullptr; }
    p.open = false;
}

void peer_release(Ctx* c) {
    PeerState& p = c->peer;
    peer_detach(p);
    if (p.local) cudaFree(p.local);
    if (p.err_host) cudaFreeHost(p.err_host);
    p = PeerState();
}

}  // namespace sfe

using namespace sfe;

extern "C" int sfe_peer_create(uint32_t rank, uint32_t world, uint64_t bytes, void* handle) {
    const char* fn = "sfe_peer_create";
    if (ensure_init(fn)) return SFE_FAILURE;
    Ctx* c = ctx();
    if (!handle) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if (world < 1 || world > SFE_PEER_MAX || rank >= world) { set_error(fn, "rank %u / world %u outside [1, %d]", rank, world, SFE_PEER_MAX); return SFE_FAILURE; }
    static_assert(sizeof(cudaIpcMemHandle_t) == SFE_PEER_HANDLE_BYTES, "handle size");
    SFE_CUDA(fn, cudaStreamSynchronize(c->stream));
    if (c->sim.loaded && c->sim.peer_mode) sim_release(c);   // its {x,y,z,m} arrays live in the window that goes away
    peer_release(c);
    PeerState& p = c->peer;
    p.rank = rank; p.world = world;
    p.bytes = (bytes + 255u) & ~(uint64_t)255u;
    SFE_CUDA(fn, cudaMalloc(&p.local, p.bytes + SFE_PEER_CTRL_BYTES));
    SFE_CUDA(fn, cudaMemset(p.local, 0, p.bytes + SFE_PEER_CTRL_BYTES));
    SFE_CUDA(fn, cudaDeviceSynchronize());   // the zeroed flags must be in place before any peer can store into them
    SFE_CUDA(fn, cudaHostAlloc((void**)&p.err_host, 64, cudaHostAllocMapped));
    *p.err_host = 0u;
    SFE_CUDA(fn, cudaHostGetDevicePointer((void**)&p.err_dev, p.err_host, 0));
    cudaIpcMemHandle_t h;
    SFE_CUDA(fn, cudaIpcGetMemHandle(&h, p.local));
    memcpy(handle, &h, sizeof h);
    p.base[rank] = p.local;
    p.created = true;
    return SFE_SUCCESS;
}

extern "C" int sfe_peer_open(const void* handles) {
    const char* fn = "sfe_peer_open";
    if (ensure_init(fn)) return SFE_FAILURE;
    PeerState& p = ctx()->peer;
    if (!p.created) { set_error(fn, "call sfe_peer_create() first"); return SFE_FAILURE; }
    if (p.open) { set_error(fn, "the window is already open"); return SFE_FAILURE; }
    if (!handles) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    for (uint32_t r = 0; r < p.world; r++) {
        if (r == p.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * SFE_PEER_HANDLE_BYTES, sizeof h);
        cudaError_t e = cudaIpcOpenMemHandle(&p.base[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error(fn, "cudaIpcOpenMemHandle(rank %u) failed: %s", r, cudaGetErrorString(e));
            for (uint32_t q = 0; q < r; q++) if (q != p.rank && p.base[q]) { cudaIpcCloseMemHandle(p.base[q]); p.base[q] = nullptr; }
            return SFE_FAILURE;
        }
    }
    p.open = true;
    return SFE_SUCCESS;
}

extern "C" void* sfe_peer_window(void) { return ctx() ? ctx()->peer.local : nullptr; }

extern "C" int sfe_dev_peer_put(uint64_t offset, uint64_t bytes, void* stream) {
    if (ensure_init("sfe_dev_peer_put")) return SFE_FAILURE;
    return launch_peer_put(ctx(), offset, bytes, (cudaStream_t)stream, "sfe_dev_peer_put");
}

extern "C" int sfe_dev_peer_barrier(void* stream) {
    if (ensure_init("sfe_dev_peer_barrier")) return SFE_FAILURE;
    return launch_peer_barrier(ctx(), (cudaStream_t)stream, "sfe_dev_peer_barrier");
}

extern "C" int sfe_peer_detach(void) {
    Ctx* c = ctx();
    if (!c) return SFE_SUCCESS;
    cudaDeviceSynchronize();
    if (c->sim.loaded && c->sim.peer_mode) sim_release(c);
    const int rc = peer_status(c, "sfe_peer_detach");
    peer_detach(c->peer);
    return rc;
}

extern "C" int sfe_peer_close(void) {
    Ctx* c = ctx();
    if (!c) return SFE_SUCCESS;
    cudaDeviceSynchronize();
    if (c->sim.loaded && c->sim.peer_mode) sim_release(c);
    const int rc = peer_status(c, "sfe_peer_close");
    peer_release(c);
    return rc;
}

// ... plus room for one frame's integer accumulation buffer behind the state blocks (sfe_sim_vis_draw() in multi-rank mode)
extern "C" uint64_t sfe_sim_peer_bytes_vis(uint32_t N, uint32_t dof, uint32_t sizex, uint32_t sizey) {
    return sfe_sim_peer_bytes(N, dof) + 16ull * (uint64_t)sizex * (uint64_t)sizey;
}

extern "C" uint64_t sfe_sim_peer_bytes(uint32_t N, uint32_t dof) {
    const uint64_t n = ((uint64_t)N + 63u) & ~(uint64_t)63u;   // keeps every array 256-byte aligned
    return 2u * 16u * n + 16u * n + (dof == 16u ? 4u * n : 0u);   // {x,y,z,m} x 2, acc + pot, rho
}
