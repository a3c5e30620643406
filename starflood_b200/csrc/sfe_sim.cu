// sfe_sim.cu - device-resident kick-drift-kick leapfrog around the force kernels (SURVEY 8f rows N1 + N2).
//
// The reference keeps the state on the host and runs the O(N) parts of sim_step() (simulation.c:348-519) and the root
// box of sim_solv() (simulation.c:245-301) as serial host loops; with the force evaluation at a few milliseconds those
// loops and the PCIe round trip of pos/mas/acc/pot become the step.  Here the state never leaves HBM between steps:
//
//   kick_drift_kernel    vel += 0.5 dt acc ; pos += dt vel                       (simulation.c:381-395)
//   root_box_kernel      mean |pos| in double, pcg4d jitter, root cube            (simulation.c:245-301, rng.h:13-40)
//   [tree build + walk | PP] [+ SPH]                                              (the engine's force kernels)
//   kick_energy_kernel   vel += 0.5 dt acc ; ken, pen ; T = sum ken, V = sum pen  (simulation.c:423-504)
//
// fp32 arithmetic is written with __fmul_rn / __fadd_rn: the reference is built in ISO C mode (no FMA contraction), and
// vel / pos / ken / pen must match it bit for bit.  The three double sums of the reference (mean radius, T, V) are serial
// Kahan folds; here they are double-double tree reductions in a fixed order, i.e. the exactly rounded sum, which a Kahan
// fold reproduces to within 2 ulp of double.  The root box is that mean cast to fp32, so it is the reference's box unless
// the mean sits within ~1e-16 (relative) of an fp32 rounding boundary.
#include "sfe_internal.cuh"

namespace sfe {

constexpr int SIM_THREADS = 256;
constexpr int SIM_MAX_BLOCKS = 1024;   // partial sums per reduction

struct __align__(16) dd { double hi, lo; };
struct TvPeers { double* slot[SFE_PEER_MAX]; uint32_t world; };   // world == 0: single rank, {T, V} go to tv[] directly

__device__ __forceinline__ dd dd_add(dd a, dd b) {
    // error-free sum of the high parts (Knuth two-sum), low parts added in, renormalised (fast two-sum)
    const double s = __dadd_rn(a.hi, b.hi);
    const double bb = __dsub_rn(s, a.hi);
    const double e = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(b.hi, bb));
    const double lo = __dadd_rn(e, __dadd_rn(a.lo, b.lo));
    const double hi = __dadd_rn(s, lo);
    return {hi, __dsub_rn(lo, __dsub_rn(hi, s))};
}
__device__ __forceinline__ dd dd_add(dd a, double x) { return dd_add(a, dd{x, 0.0}); }

__device__ __forceinline__ dd dd_shfl_down(dd v, int delta) {
    return {__shfl_down_sync(0xFFFFFFFFu, v.hi, delta), __shfl_down_sync(0xFFFFFFFFu, v.lo, delta)};
}

// Block-wide double-double sums of NV values per thread in a fixed order; valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(dd (&v)[NV], dd (*sh)[SIM_THREADS / 32]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[k] = dd_add(v[k], dd_shfl_down(v[k], d));
        if (lane == 0) sh[k][warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            dd s = sh[k][0];
            for (int w = 1; w < SIM_THREADS / 32; w++) s = dd_add(s, sh[k][w]);
            v[k] = s;
        }
    }
}

// Last-block-done: the block that draws the final ticket folds the per-block partials in block order.
template <int NV>
__device__ __forceinline__ bool grid_sum(dd (&v)[NV], dd* __restrict__ partial, uint32_t* __restrict__ ticket) {
    __shared__ dd sh[NV][SIM_THREADS / 32];
    __shared__ bool last;
    block_sum<NV>(v, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) partial[(size_t)k * SIM_MAX_BLOCKS + blockIdx.x] = v[k];
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    // fold gridDim.x partials: one warp, lane-strided, then the same shuffle tree (fixed order for a given grid)
    if (threadIdx.x < 32) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            dd s{0.0, 0.0};
            for (uint32_t b = threadIdx.x; b < gridDim.x; b += 32u) {
                const double2 q = __ldcg(reinterpret_cast<const double2*>(partial + (size_t)k * SIM_MAX_BLOCKS + b));
                s = dd_add(s, dd{q.x, q.y});
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) s = dd_add(s, dd_shfl_down(s, d));
            v[k] = s;
        }
    }
    if (threadIdx.x == 0) *ticket = 0u;   // ready for the next launch on this stream
    return threadIdx.x == 0;
}

// ---------------------------------------------------------------- kick_0 + drift (simulation.c:381-395)
// vel[i] += (real)0.5 * dt * acc[i]  ->  ((0.5 * dt) * acc) added to vel ; pos[i] += dt * vel[i]
__global__ void __launch_bounds__(SIM_THREADS) kick_drift_kernel(float4* __restrict__ pm4, float* __restrict__ vel,
                                                                 const float* __restrict__ acc, uint32_t i0, uint32_t ni,
                                                                 float dt) {
    const float hdt = __fmul_rn(0.5f, dt);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ni; t += gridDim.x * blockDim.x) {
        const uint32_t i = i0 + t;
        float4 p = pm4[i];
        float v[3], q[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            v[k] = __fadd_rn(vel[3u * i + k], __fmul_rn(hdt, acc[3u * i + k]));
            q[k] = __fadd_rn(q[k], __fmul_rn(dt, v[k]));
            vel[3u * i + k] = v[k];
        }
        pm4[i] = make_float4(q[0], q[1], q[2], p.w);
    }
}

// The same for a rank's target slice in multi-rank mode, fused with the all-gather that SURVEY 8(e) puts after each drift:
// the moved {x,y,z,m} record goes straight into the NEXT array of every rank's window (16-byte stores over NVLink peer
// pointers), so no separate collective runs; the arrival flags follow in peer_barrier_kernel (sfe_peer.cu).
__global__ void __launch_bounds__(SIM_THREADS) kick_drift_put_kernel(PeerPtrs pp, uint64_t off_cur, uint64_t off_nxt,
                                                                     float* __restrict__ vel, const float* __restrict__ acc,
                                                                     uint32_t i0, uint32_t ni, float dt) {
    const float hdt = __fmul_rn(0.5f, dt);
    const float4* __restrict__ cur = reinterpret_cast<const float4*>(pp.base[pp.rank] + off_cur);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ni; t += gridDim.x * blockDim.x) {
        const uint32_t i = i0 + t;
        const float4 p = cur[i];
        float v[3], q[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            v[k] = __fadd_rn(vel[3u * i + k], __fmul_rn(hdt, acc[3u * i + k]));
            q[k] = __fadd_rn(q[k], __fmul_rn(dt, v[k]));
            vel[3u * i + k] = v[k];
        }
        const float4 out = make_float4(q[0], q[1], q[2], p.w);
        for (uint32_t r = 0; r < pp.world; r++) reinterpret_cast<float4*>(pp.base[r] + off_nxt)[i] = out;
    }
}

// ---------------------------------------------------------------- root box (simulation.c:245-301)
__device__ __forceinline__ void pcg4d_dev(uint32_t (&v)[4]) {   // rng.h:13-40
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = v[k] * 0x0019660Du + 0x3C6EF35Fu;
    v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] ^= v[k] >> 16;
    v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
}

struct BoxOut { float bounds[6]; uint32_t pad; uint32_t seq; double rbar; };   // seq is written last

__global__ void __launch_bounds__(SIM_THREADS) root_box_kernel(const float4* __restrict__ pm4, uint32_t N,
                                                               uint32_t step_number, dd* __restrict__ partial,
                                                               uint32_t* __restrict__ ticket, BoxOut* __restrict__ out_dev,
                                                               volatile BoxOut* __restrict__ out_host, uint32_t seq) {
    dd v[1] = {{0.0, 0.0}};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float4 p = pm4[i];
        const double x = (double)p.x, y = (double)p.y, z = (double)p.z;
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
        v[0] = dd_add(v[0], __dsqrt_rn(r2));
    }
    if (!grid_sum<1>(v, partial, ticket)) return;
    const double rbar = __dmul_rn(v[0].hi, __ddiv_rn(1.0, (double)N));   // average_radius *= 1.0 / (double)sim.N
    uint32_t s[4] = {0x638FE17Bu + step_number, 0xB10BDCE7u, 0x7B8ADB9Bu, 0xD8C5850Eu};
    pcg4d_dev(s);
    const double inv = 2.3283062977608182109179324470460414886474609375e-10;   // INV_PCG32_MAX (rng.h:9)
    float b[6];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double jit = __dsub_rn(__dmul_rn(2.0, __dmul_rn(inv, (double)s[k])), 1.0);
        const double off = __dmul_rn(__dmul_rn(0.5, jit), rbar);
        b[k] = (float)__dadd_rn(__dmul_rn(-2.0, rbar), off);
        b[3 + k] = (float)__dadd_rn(__dmul_rn(2.0, rbar), off);
    }
#pragma unroll
    for (int k = 0; k < 6; k++) { out_dev->bounds[k] = b[k]; out_host->bounds[k] = b[k]; }
    out_dev->rbar = rbar; out_host->rbar = rbar;
    __threadfence_system();
    out_host->seq = seq;
}

// ---------------------------------------------------------------- kick_1 + energies (simulation.c:423-504)
__global__ void __launch_bounds__(SIM_THREADS) kick_energy_kernel(const float4* __restrict__ pm4, float* __restrict__ vel,
                                                                  const float* __restrict__ acc, const float* __restrict__ pot,
                                                                  float* __restrict__ ken, float* __restrict__ pen,
                                                                  uint32_t i0, uint32_t ni, float dt, int do_kick,
                                                                  dd* __restrict__ partial, uint32_t* __restrict__ ticket,
                                                                  double* __restrict__ tv, TvPeers tp) {
    const float hdt = __fmul_rn(0.5f, dt);
    dd s[2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ni; t += gridDim.x * blockDim.x) {
        const uint32_t i = i0 + t;
        const float m = pm4[i].w;
        float v[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            v[k] = vel[3u * i + k];
            if (do_kick) { v[k] = __fadd_rn(v[k], __fmul_rn(hdt, acc[3u * i + k])); vel[3u * i + k] = v[k]; }
        }
        const float v2 = __fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2]));
        const float ke = __fmul_rn(__fmul_rn(0.5f, m), v2);   // ken[i] = (real)0.5 * m_i * v2
        const float pe = __fmul_rn(m, pot[i]);                // pen[i] = mas[i] * pot[i]
        ken[i] = ke; pen[i] = pe;
        s[0] = dd_add(s[0], (double)ke);
        s[1] = dd_add(s[1], (double)pe);
    }
    if (!grid_sum<2>(s, partial, ticket)) return;
    if (tp.world == 0u) { tv[0] = s[0].hi; tv[1] = s[1].hi; return; }
    // multi-rank: the slice's double-double sums go into slot [rank] of every rank's control page; tv_fold_kernel adds
    // the slots in rank order after the barrier
    for (uint32_t r = 0; r < tp.world; r++) {
        double* d = tp.slot[r];
        d[0] = s[0].hi; d[1] = s[0].lo; d[2] = s[1].hi; d[3] = s[1].lo;
    }
}

__global__ void tv_fold_kernel(const double* __restrict__ slots /* [world][4] */, uint32_t world, double* __restrict__ tv) {
    dd T{0.0, 0.0}, V{0.0, 0.0};
    for (uint32_t r = 0; r < world; r++) {
        T = dd_add(T, dd{__ldcg(&slots[4u * r]), __ldcg(&slots[4u * r + 1u])});
        V = dd_add(V, dd{__ldcg(&slots[4u * r + 2u]), __ldcg(&slots[4u * r + 3u])});
    }
    tv[0] = T.hi; tv[1] = V.hi;
}

__global__ void unpack_pos_kernel(const float4* __restrict__ pm4, float* __restrict__ pos, float* __restrict__ mas, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pm4[i];
    if (pos) { pos[3u * i] = p.x; pos[3u * i + 1u] = p.y; pos[3u * i + 2u] = p.z; }
    if (mas) mas[i] = p.w;
}

// ---------------------------------------------------------------------------- launchers
static uint32_t sim_grid(Ctx* c, uint32_t n) {
    uint32_t g = (n + SIM_THREADS - 1) / SIM_THREADS, cap = (uint32_t)c->sm_count * 4u;
    if (cap > SIM_MAX_BLOCKS) cap = SIM_MAX_BLOCKS;
    return g < 1u ? 1u : (g < cap ? g : cap);
}

static int sim_scratch(Ctx* c, const char* fn, cudaStream_t s) {
    SimState& st = c->sim;
    if (st.scratch_ready) return SFE_SUCCESS;
    if (st.partial.reserve(sizeof(dd) * 2 * SIM_MAX_BLOCKS, fn) || st.ticket.reserve(16, fn) || st.box_dev.reserve(sizeof(BoxOut), fn) ||
        st.tv_dev.reserve(sizeof(double) * 2, fn))
        return SFE_FAILURE;
    SFE_CUDA(fn, cudaMemsetAsync(st.ticket.p, 0, 16, s));
    SFE_CUDA(fn, cudaHostAlloc(&st.box_host, sizeof(BoxOut), cudaHostAllocMapped));
    memset(st.box_host, 0, sizeof(BoxOut));
    SFE_CUDA(fn, cudaHostGetDevicePointer(&st.box_host_dev, st.box_host, 0));
    SFE_CUDA(fn, cudaMallocHost(&st.tv_host, sizeof(double) * 2 * SFE_SIM_TV_RING));
    st.scratch_ready = true;
    return SFE_SUCCESS;
}

int launch_kick_drift(Ctx* c, float4* pm4, float* vel, const float* acc, uint32_t i0, uint32_t ni, float dt, cudaStream_t s) {
    if (ni == 0) return SFE_SUCCESS;
    kick_drift_kernel<<<sim_grid(c, ni), SIM_THREADS, 0, s>>>(pm4, vel, acc, i0, ni, dt);
    return check_launch("sfe_dev_kick_drift");
}

int launch_root_box_dev(Ctx* c, const float4* pm4, uint32_t N, uint32_t step_number, const float** bounds_dev, cudaStream_t s) {
    const char* fn = "sfe_dev_root_box";
    if (N == 0) { set_error(fn, "N must be at least 1 (the reference divides by N, simulation.c:277)"); return SFE_FAILURE; }
    if (sim_scratch(c, fn, s)) return SFE_FAILURE;
    SimState& st = c->sim;
    const uint32_t seq = ++st.box_seq ? st.box_seq : ++st.box_seq;
    root_box_kernel<<<sim_grid(c, N), SIM_THREADS, 0, s>>>(pm4, N, step_number, st.partial.as<dd>(), st.ticket.as<uint32_t>(),
                                                           st.box_dev.as<BoxOut>(), (volatile BoxOut*)st.box_host_dev, seq);
    if (check_launch(fn)) return SFE_FAILURE;
    *bounds_dev = st.box_dev.as<BoxOut>()->bounds;
    return SFE_SUCCESS;
}

// the box of the last launch_root_box_dev() on the host: the device writes it into mapped host memory; spin on its sequence
// number (no stream operation)
int root_box_host(Ctx* c, float bounds[6], cudaStream_t s, const char* fn) {
    SimState& st = c->sim;
    const uint32_t seq = st.box_seq;
    volatile BoxOut* h = (volatile BoxOut*)st.box_host;
    uint64_t spins = 0;
    while (h->seq != seq) {
        if ((++spins & 0xFFFFu) == 0u) {
            cudaError_t q = cudaStreamQuery(s);
            if (q != cudaErrorNotReady && h->seq != seq) {
                set_error(fn, q == cudaSuccess ? "the root box never arrived" : cudaGetErrorString(q));
                return SFE_FAILURE;
            }
        }
    }
    for (int k = 0; k < 6; k++) bounds[k] = h->bounds[k];
    return SFE_SUCCESS;
}

int launch_root_box(Ctx* c, const float4* pm4, uint32_t N, uint32_t step_number, float bounds[6], cudaStream_t s) {
    const float* dev = nullptr;
    if (launch_root_box_dev(c, pm4, N, step_number, &dev, s)) return SFE_FAILURE;
    return root_box_host(c, bounds, s, "sfe_dev_root_box");
}

// Mean |pos| left in device memory (visualization.c:274-307 computes the same double Kahan mean as simulation.c:245-277):
// the root-box kernel without the host hand-over.  *rbar_dev stays valid until the next root box / mean radius.
int launch_mean_radius(Ctx* c, const float4* pm4, uint32_t N, const double** rbar_dev, cudaStream_t s) {
    const char* fn = "sfe_vis_draw";
    if (N == 0) { set_error(fn, "N must be at least 1 (the reference divides by N, visualization.c:304)"); return SFE_FAILURE; }
    if (sim_scratch(c, fn, s)) return SFE_FAILURE;
    SimState& st = c->sim;
    const uint32_t seq = ++st.box_seq ? st.box_seq : ++st.box_seq;
    root_box_kernel<<<sim_grid(c, N), SIM_THREADS, 0, s>>>(pm4, N, 0u, st.partial.as<dd>(), st.ticket.as<uint32_t>(),
                                                           st.box_dev.as<BoxOut>(), (volatile BoxOut*)st.box_host_dev, seq);
    *rbar_dev = &st.box_dev.as<BoxOut>()->rbar;
    return check_launch(fn);
}

// the resident state's particle records and potentials (single-rank mode), for the splat of sfe_sim_vis_draw()
static float4* sim_pm4(Ctx* c);
static float* sim_pot(Ctx* c);

// Views of the resident state for vis_draw().  In multi-rank mode every rank holds all positions but only the potentials
// of its own slice [i0, i0+ni): it splats that slice, and the accumulation buffers are summed over the windows (the frame's
// region starts at *vis_off, behind the state blocks).
int sim_resident_views(Ctx* c, const float4** pm4, const float** pot, uint32_t* N, const char* fn,
                       uint32_t* i0, uint32_t* ni, uint64_t* vis_off, bool* sliced) {
    SimState& st = c->sim;
    if (!st.loaded) { set_error(fn, "no resident state: call sfe_sim_upload() first"); return SFE_FAILURE; }
    *pm4 = sim_pm4(c); *pot = sim_pot(c); *N = st.N;
    *sliced = st.peer_mode; *i0 = st.i0; *ni = st.ni;
    *vis_off = sfe_sim_peer_bytes(st.N, st.dof);
    return SFE_SUCCESS;
}

int launch_kick_energy(Ctx* c, const float4* pm4, float* vel, const float* acc, const float* pot, float* ken, float* pen,
                       uint32_t i0, uint32_t ni, float dt, int do_kick, double* tv_dev, cudaStream_t s) {
    const char* fn = "sfe_dev_kick_energy";
    if (sim_scratch(c, fn, s)) return SFE_FAILURE;
    SimState& st = c->sim;
    if (ni == 0) { SFE_CUDA(fn, cudaMemsetAsync(tv_dev, 0, 2 * sizeof(double), s)); return SFE_SUCCESS; }
    TvPeers tp; tp.world = 0u;
    kick_energy_kernel<<<sim_grid(c, ni), SIM_THREADS, 0, s>>>(pm4, vel, acc, pot, ken, pen, i0, ni, dt, do_kick,
                                                              st.partial.as<dd>(), st.ticket.as<uint32_t>() + 1, tv_dev, tp);
    return check_launch(fn);
}

// Multi-rank kick_1 + energies: the slice's sums are stored into every rank's control page, the ranks exchange arrival
// flags, and each of them folds the slots in rank order (the same {T, V} everywhere).  An empty slice still takes part.
static int launch_kick_energy_peers(Ctx* c, const float4* pm4, float dt, cudaStream_t s, const char* fn) {
    SimState& st = c->sim;
    PeerState& pr = c->peer;
    const uint32_t parity = st.tv_parity; st.tv_parity ^= 1u;
    TvPeers tp; tp.world = pr.world;
    for (uint32_t r = 0; r < pr.world; r++)
        tp.slot[r] = reinterpret_cast<PeerCtrl*>((char*)pr.base[r] + pr.bytes)->tv[parity][pr.rank];
    const float* acc = reinterpret_cast<const float*>((const char*)pr.local + st.acc_off);
    const float* pot = reinterpret_cast<const float*>((const char*)pr.local + st.pot_off);
    kick_energy_kernel<<<sim_grid(c, st.ni), SIM_THREADS, 0, s>>>(pm4, st.vel.as<float>(), acc, pot,
                                                                 st.ken.as<float>(), st.pen.as<float>(), st.i0, st.ni, dt, 1,
                                                                 st.partial.as<dd>(), st.ticket.as<uint32_t>() + 1,
                                                                 st.tv_dev.as<double>(), tp);
    if (check_launch(fn)) return SFE_FAILURE;
    if (launch_peer_barrier(c, s, fn)) return SFE_FAILURE;
    tv_fold_kernel<<<1, 1, 0, s>>>(&reinterpret_cast<PeerCtrl*>((char*)pr.local + pr.bytes)->tv[parity][0][0], pr.world,
                                   st.tv_dev.as<double>());
    return check_launch(fn);
}

static int launch_kick_drift_peers(Ctx* c, float dt, cudaStream_t s, const char* fn) {
    SimState& st = c->sim;
    if (st.ni) {
        const float* acc = reinterpret_cast<const float*>((const char*)c->peer.local + st.acc_off);
        kick_drift_put_kernel<<<sim_grid(c, st.ni), SIM_THREADS, 0, s>>>(peer_ptrs(c), st.pm4_off[st.cur], st.pm4_off[st.cur ^ 1],
                                                                        st.vel.as<float>(), acc, st.i0, st.ni, dt);
        if (check_launch(fn)) return SFE_FAILURE;
    }
    if (launch_peer_barrier(c, s, fn)) return SFE_FAILURE;
    st.cur ^= 1;
    return SFE_SUCCESS;
}

static float4* sim_pm4(Ctx* c) {
    SimState& st = c->sim;
    return st.peer_mode ? reinterpret_cast<float4*>((char*)c->peer.local + st.pm4_off[st.cur]) : st.pm4.as<float4>();
}
static float* sim_acc(Ctx* c) {
    SimState& st = c->sim;
    return st.peer_mode ? reinterpret_cast<float*>((char*)c->peer.local + st.acc_off) : st.acc.as<float>();
}
static float* sim_pot(Ctx* c) {
    SimState& st = c->sim;
    return st.peer_mode ? reinterpret_cast<float*>((char*)c->peer.local + st.pot_off) : st.pot.as<float>();
}
static float* sim_rho(Ctx* c) {
    SimState& st = c->sim;
    return st.peer_mode ? reinterpret_cast<float*>((char*)c->peer.local + st.rho_off) : st.rho.as<float>();
}

int launch_unpack(const float4* pm4, float* pos, float* mas, uint32_t n, cudaStream_t s) {
    if (n == 0) return SFE_SUCCESS;
    unpack_pos_kernel<<<(n + 255) / 256, 256, 0, s>>>(pm4, pos, mas, n);
    return check_launch("sfe_sim_download");
}

void sim_release(Ctx* c) {
    SimState& st = c->sim;
    DevBuf* bufs[] = {&st.pm4, &st.vel, &st.acc, &st.pot, &st.ken, &st.pen, &st.rad, &st.rho, &st.prs, &st.stage,
                      &st.partial, &st.ticket, &st.box_dev, &st.tv_dev};
    for (DevBuf* b : bufs) b->release();
    if (st.box_host) cudaFreeHost(st.box_host);
    if (st.tv_host) cudaFreeHost(st.tv_host);
    st = SimState();
}

}  // namespace sfe

using namespace sfe;

// ============================================================================ device entry points
extern "C" int sfe_dev_kick_drift(void* pm4, float* vel, const float* acc, uint32_t i0, uint32_t ni, float dt, void* stream) {
    if (ensure_init("sfe_dev_kick_drift")) return SFE_FAILURE;
    return launch_kick_drift(ctx(), (float4*)pm4, vel, acc, i0, ni, dt, (cudaStream_t)stream);
}

extern "C" int sfe_dev_root_box(const void* pm4, uint32_t N, uint32_t step_number, float bounds[6], void* stream) {
    if (ensure_init("sfe_dev_root_box")) return SFE_FAILURE;
    if (!bounds) { set_error("sfe_dev_root_box", "null pointer argument"); return SFE_FAILURE; }
    return launch_root_box(ctx(), (const float4*)pm4, N, step_number, bounds, (cudaStream_t)stream);
}

extern "C" int sfe_dev_kick_energy(const void* pm4, float* vel, const float* acc, const float* pot, float* ken, float* pen,
                                   uint32_t i0, uint32_t ni, float dt, int do_kick, double* tv, void* stream) {
    if (ensure_init("sfe_dev_kick_energy")) return SFE_FAILURE;
    return launch_kick_energy(ctx(), (const float4*)pm4, vel, acc, pot, ken, pen, i0, ni, dt, do_kick, tv, (cudaStream_t)stream);
}

// ============================================================================ host entry points: resident state
static int sim_have(const char* fn, uint32_t need_dof = 0) {
    if (ensure_init(fn)) return SFE_FAILURE;
    SimState& st = ctx()->sim;
    if (!st.loaded) { set_error(fn, "no resident state: call sfe_sim_upload() first"); return SFE_FAILURE; }
    if (need_dof && st.dof < need_dof) { set_error(fn, "the resident state has SIM_DOF=%u; the SPH blocks need 16", st.dof); return SFE_FAILURE; }
    return SFE_SUCCESS;
}

extern "C" int sfe_sim_upload(const float* mem, uint32_t N, uint32_t dof) {
    const char* fn = "sfe_sim_upload";
    if (ensure_init(fn)) return SFE_FAILURE;
    if (!mem && N) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if (dof != 13u && dof != 16u) { set_error(fn, "SIM_DOF must be 13 or 16 (simulation.h:40-46), got %u", dof); return SFE_FAILURE; }
    if (N == 0) { set_error(fn, "N must be at least 1"); return SFE_FAILURE; }
    Ctx* c = ctx();
    SimState& st = c->sim;
    cudaStream_t s = c->stream;
    const size_t n = N;
    if (sim_scratch(c, fn, s)) return SFE_FAILURE;
    // multi-rank mode: an open peer window with more than one rank; the particles are partitioned by target slice
    const PeerState& pr = c->peer;
    st.peer_mode = pr.open && pr.world > 1u;
    st.i0 = 0u; st.ni = N; st.cur = 0; st.tv_parity = 0u;
    if (st.peer_mode) {
        if (pr.bytes < sfe_sim_peer_bytes(N, dof)) {
            set_error(fn, "the peer window holds %llu bytes; N=%u SIM_DOF=%u needs sfe_sim_peer_bytes() = %llu",
                      (unsigned long long)pr.bytes, N, dof, (unsigned long long)sfe_sim_peer_bytes(N, dof));
            st.peer_mode = false; st.loaded = false;
            return SFE_FAILURE;
        }
        const uint64_t n64 = ((uint64_t)N + 63u) & ~(uint64_t)63u;
        st.pm4_off[0] = 0u; st.pm4_off[1] = 16u * n64; st.acc_off = 32u * n64; st.pot_off = 44u * n64; st.rho_off = 48u * n64;
        st.i0 = (uint32_t)(((uint64_t)N * pr.rank) / pr.world);
        st.ni = (uint32_t)(((uint64_t)N * (pr.rank + 1u)) / pr.world) - st.i0;
    }
    if ((!st.peer_mode && (st.pm4.reserve(n * 16, fn) || st.acc.reserve(n * 12, fn) || st.pot.reserve(n * 4, fn))) || st.vel.reserve(n * 12, fn) ||
        st.ken.reserve(n * 4, fn) || st.pen.reserve(n * 4, fn) || st.stage.reserve(n * 16, fn))
        return SFE_FAILURE;
    if (dof == 16u && (st.rad.reserve(n * 4, fn) || (!st.peer_mode && st.rho.reserve(n * 4, fn)) || st.prs.reserve(n * 4, fn))) return SFE_FAILURE;
    st.loaded = false;
    // RAW layout: block p starts at mem + N*p (simulation.c:52); pos/mas go through the staging buffer into pm4
    float* stage = st.stage.as<float>();
    if (h2d(fn, stage, mem, n * 12, s) || h2d(fn, stage + 3 * n, mem + 9 * n, n * 4, s)) return SFE_FAILURE;
    if (launch_pack_pm4(stage, stage + 3 * n, sim_pm4(c), N, s)) return SFE_FAILURE;
    if (h2d(fn, st.vel.p, mem + 3 * n, n * 12, s) || h2d(fn, sim_acc(c), mem + 6 * n, n * 12, s) ||
        h2d(fn, sim_pot(c), mem + 10 * n, n * 4, s) || h2d(fn, st.ken.p, mem + 11 * n, n * 4, s) ||
        h2d(fn, st.pen.p, mem + 12 * n, n * 4, s))
        return SFE_FAILURE;
    if (dof == 16u && (h2d(fn, st.rad.p, mem + 13 * n, n * 4, s) || h2d(fn, sim_rho(c), mem + 14 * n, n * 4, s) ||
                       h2d(fn, st.prs.p, mem + 15 * n, n * 4, s)))
        return SFE_FAILURE;
    // multi-rank: nobody may store into a window (the first rho exchange, the first drift) before its owner has finished
    // filling it, and nobody may refill a window that a slower peer still reads: one exchange of flags
    if (st.peer_mode && launch_peer_barrier(c, s, fn)) return SFE_FAILURE;
    SFE_CUDA(fn, cudaStreamSynchronize(s));
    if (st.peer_mode && peer_status(c, fn)) return SFE_FAILURE;
    st.N = N; st.dof = dof; st.loaded = true;
    return SFE_SUCCESS;
}

extern "C" int sfe_sim_download(float* mem, uint32_t N, uint32_t dof, uint32_t blocks) {
    const char* fn = "sfe_sim_download";
    if (sim_have(fn)) return SFE_FAILURE;
    Ctx* c = ctx();
    SimState& st = c->sim;
    if (!mem) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if (N != st.N || dof != st.dof) { set_error(fn, "resident state is N=%u SIM_DOF=%u, asked for N=%u SIM_DOF=%u", st.N, st.dof, N, dof); return SFE_FAILURE; }
    cudaStream_t s = c->stream;
    const size_t n = N;
    if (blocks & (SFE_SIM_POS | SFE_SIM_MAS)) {
        float* stage = st.stage.as<float>();
        if (launch_unpack(sim_pm4(c), (blocks & SFE_SIM_POS) ? stage : nullptr, (blocks & SFE_SIM_MAS) ? stage + 3 * n : nullptr, N, s))
            return SFE_FAILURE;
        if ((blocks & SFE_SIM_POS) && d2h(fn, mem, stage, n * 12, s)) return SFE_FAILURE;
        if ((blocks & SFE_SIM_MAS) && d2h(fn, mem + 9 * n, stage + 3 * n, n * 4, s)) return SFE_FAILURE;
    }
    // per-particle blocks that only the owning rank maintains: the slice [i0, i0+ni) (everything on a single rank)
    const size_t a = st.i0, m = st.ni;
    if ((blocks & SFE_SIM_VEL) && d2h(fn, mem + 3 * n + 3 * a, st.vel.as<float>() + 3 * a, m * 12, s)) return SFE_FAILURE;
    if ((blocks & SFE_SIM_ACC) && d2h(fn, mem + 6 * n + 3 * a, sim_acc(c) + 3 * a, m * 12, s)) return SFE_FAILURE;
    if ((blocks & SFE_SIM_POT) && d2h(fn, mem + 10 * n + a, sim_pot(c) + a, m * 4, s)) return SFE_FAILURE;
    if ((blocks & SFE_SIM_KEN) && d2h(fn, mem + 11 * n + a, st.ken.as<float>() + a, m * 4, s)) return SFE_FAILURE;
    if ((blocks & SFE_SIM_PEN) && d2h(fn, mem + 12 * n + a, st.pen.as<float>() + a, m * 4, s)) return SFE_FAILURE;
    if (dof == 16u) {
        if ((blocks & SFE_SIM_RAD) && d2h(fn, mem + 13 * n, st.rad.p, n * 4, s)) return SFE_FAILURE;
        if ((blocks & SFE_SIM_RHO) && d2h(fn, mem + 14 * n, sim_rho(c), n * 4, s)) return SFE_FAILURE;
        if ((blocks & SFE_SIM_PRS) && d2h(fn, mem + 15 * n, st.prs.p, n * 4, s)) return SFE_FAILURE;
    }
    SFE_CUDA(fn, cudaStreamSynchronize(s));
    return SFE_SUCCESS;
}

// sim_solv() (simulation.c:231-345) on the resident state; enqueues only (the tree path waits for the root box)
static int sim_solv_enqueue(Ctx* c, const sfe_sim_params_t* p, uint32_t step_number, const char* fn) {
    SimState& st = c->sim;
    cudaStream_t s = c->stream;
    const float4* pm4 = sim_pm4(c);
    const uint32_t i0 = st.i0, ni = st.ni;
    if (p->solver & SFE_SOLVER_TREE) {
        float box[6];
        // the box stays on the device: the build reads it there, the host copy (sfe_sim_last_box) is picked up behind the
        // status check below
        const float* box_dev = nullptr;
        if (launch_root_box_dev(c, pm4, st.N, step_number, &box_dev, s)) return SFE_FAILURE;
        // the reference ignores tree_build()'s return value (simulation.c:312) and walks whatever was built; the
        // engine raises the failure instead (a truncated tree has no defined contents here)
        if (launch_tree_build(c, box, pm4, st.N, p->depth_max, p->nodes_max, s, box_dev)) return SFE_FAILURE;
        if (tree_status(c, s)) return SFE_FAILURE;
        if (root_box_host(c, box, s, fn)) return SFE_FAILURE;
        for (int k = 0; k < 6; k++) st.last_box[k] = box[k];
        if (st.peer_mode) {
            // every rank walks a slot range of the (identical) tree and delivers acc / pot to the particles' owners
            if (launch_bh_routed(c, pm4, st.N, st.acc_off, st.pot_off, p->grav_g, p->eps2, p->theta, p->flavour, s)) return SFE_FAILURE;
        } else if (launch_bh(c, pm4, st.N, i0, ni, sim_acc(c), sim_pot(c), p->grav_g, p->eps2, p->theta, p->flavour, s))
            return SFE_FAILURE;
    } else {
        if (launch_pp(c, pm4, st.N, i0, ni, sim_acc(c), sim_pot(c), p->grav_g, p->eps2, s)) return SFE_FAILURE;
    }
    if (p->solver & SFE_SOLVER_SPH) {
        if (st.dof < 16u) { set_error(fn, "SFE_SOLVER_SPH needs a SIM_DOF=16 state"); return SFE_FAILURE; }
        float* rho = sim_rho(c);
        if (st.peer_mode) {
            // Second exchange of SURVEY 8(e).  Cell-list pass: every rank evaluates a dense range of the cell-sorted
            // targets and stores rho into EVERY rank's window (the force pass needs rho of all sources), then adds the
            // pressure force into the owner's acc over NVLink.  All-pairs pass (small N; the choice is made on the device,
            // the same on every rank): the index slice, locally, and the slice is put to the peers - a no-op re-send when
            // the cell pass ran.
            SphRoute rt;
            for (uint32_t r = 0; r < SFE_PEER_MAX; r++) rt.base[r] = (char*)c->peer.base[r];
            rt.world = c->peer.world; rt.rank = c->peer.rank;
            rt.off = st.rho_off;
            if (launch_sph_density(c, pm4, st.rad.as<float>(), st.N, i0, ni, rho, s, &rt)) return SFE_FAILURE;
            if (launch_peer_barrier(c, s, fn)) return SFE_FAILURE;
            if (launch_peer_put(c, st.rho_off + 4u * (uint64_t)i0, 4u * (uint64_t)ni, s, fn)) return SFE_FAILURE;
            if (launch_sph_eos(rho, st.prs.as<float>(), st.N, p->eos_k, p->eos_n, s)) return SFE_FAILURE;
            rt.off = st.acc_off;
            if (launch_sph_force(c, pm4, st.rad.as<float>(), rho, st.prs.as<float>(), st.N, i0, ni, p->cutoff2, sim_acc(c), s,
                                 /*reuse_cells=*/true, &rt))
                return SFE_FAILURE;
            if (launch_peer_barrier(c, s, fn)) return SFE_FAILURE;
        } else {
            if (launch_sph_density(c, pm4, st.rad.as<float>(), st.N, i0, ni, rho, s)) return SFE_FAILURE;
            if (launch_sph_eos(rho, st.prs.as<float>(), st.N, p->eos_k, p->eos_n, s)) return SFE_FAILURE;
            if (launch_sph_force(c, pm4, st.rad.as<float>(), rho, st.prs.as<float>(), st.N, i0, ni, p->cutoff2, sim_acc(c), s,
                                 /*reuse_cells=*/true))
                return SFE_FAILURE;
        }
    }
    return SFE_SUCCESS;
}

static int sim_finish(Ctx* c, const sfe_sim_params_t* p, const char* fn) {
    SFE_CUDA(fn, cudaStreamSynchronize(c->stream));
    if (c->sim.peer_mode && peer_status(c, fn)) return SFE_FAILURE;
    if ((p->solver & SFE_SOLVER_TREE) && tree_status(c, c->stream)) return SFE_FAILURE;
    return SFE_SUCCESS;
}

extern "C" int sfe_sim_solv(const sfe_sim_params_t* p, uint32_t step_number) {
    const char* fn = "sfe_sim_solv";
    if (sim_have(fn)) return SFE_FAILURE;
    if (!p) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if (sim_solv_enqueue(ctx(), p, step_number, fn)) return SFE_FAILURE;
    return sim_finish(ctx(), p, fn);
}

extern "C" int sfe_sim_run(const sfe_sim_params_t* p, uint32_t first_step, uint32_t nsteps, double* energy) {
    const char* fn = "sfe_sim_step";
    if (sim_have(fn)) return SFE_FAILURE;
    if (!p) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    Ctx* c = ctx();
    SimState& st = c->sim;
    cudaStream_t s = c->stream;
    uint32_t pending = 0;   // steps whose {T, V} sit in the pinned ring and have not been handed out yet
    auto flush = [&](uint32_t upto) -> int {
        SFE_CUDA(fn, cudaStreamSynchronize(s));
        if (st.peer_mode && peer_status(c, fn)) return SFE_FAILURE;
        if ((p->solver & SFE_SOLVER_TREE) && tree_status(c, s)) return SFE_FAILURE;
        for (; pending < upto; pending++) {
            const double T = st.tv_host[2 * (pending % SFE_SIM_TV_RING)], V = st.tv_host[2 * (pending % SFE_SIM_TV_RING) + 1];
            if (energy) { energy[3 * pending] = V + T; energy[3 * pending + 1] = V; energy[3 * pending + 2] = T; }
        }
        return SFE_SUCCESS;
    };
    for (uint32_t k = 0; k < nsteps; k++) {
        if (k - pending == SFE_SIM_TV_RING && flush(k)) return SFE_FAILURE;
        if (st.peer_mode) {
            if (launch_kick_drift_peers(c, p->dt, s, fn)) return SFE_FAILURE;
            if (sim_solv_enqueue(c, p, first_step + k, fn)) return SFE_FAILURE;
            if (launch_kick_energy_peers(c, sim_pm4(c), p->dt, s, fn)) return SFE_FAILURE;
        } else {
            if (launch_kick_drift(c, st.pm4.as<float4>(), st.vel.as<float>(), st.acc.as<float>(), 0, st.N, p->dt, s)) return SFE_FAILURE;
            if (sim_solv_enqueue(c, p, first_step + k, fn)) return SFE_FAILURE;
            if (launch_kick_energy(c, st.pm4.as<float4>(), st.vel.as<float>(), st.acc.as<float>(), st.pot.as<float>(), st.ken.as<float>(),
                                   st.pen.as<float>(), 0, st.N, p->dt, 1, st.tv_dev.as<double>(), s))
                return SFE_FAILURE;
        }
        SFE_CUDA(fn, cudaMemcpyAsync(st.tv_host + 2 * (k % SFE_SIM_TV_RING), st.tv_dev.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    return flush(nsteps);
}

extern "C" int sfe_sim_step(const sfe_sim_params_t* p, uint32_t step_number, double energy[3]) {
    return sfe_sim_run(p, step_number, 1, energy);
}

extern "C" int sfe_sim_last_box(float bounds[6]) {
    if (sim_have("sfe_sim_last_box")) return SFE_FAILURE;
    for (int k = 0; k < 6; k++) bounds[k] = ctx()->sim.last_box[k];
    return SFE_SUCCESS;
}

extern "C" int sfe_sim_release(void) {
    if (!ctx()) return SFE_SUCCESS;
    cudaStreamSynchronize(ctx()->stream);
    sim_release(ctx());
    return SFE_SUCCESS;
}
