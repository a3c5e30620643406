// sfe_api.cu - the C ABI (include/sfengine.h): context lifecycle, host <-> device staging and the
// host / device entry points that the reference-side shims and the Python host binding call.
#include "sfe_internal.cuh"
#include <stdlib.h>
#include <map>
#include <mutex>

namespace sfe {

static Ctx* g_ctx = nullptr;
static char g_err[512] = "";
static std::map<void*, uint64_t> g_registered;

Ctx* ctx() { return g_ctx; }

void set_error(const char* fn, const char* fmt, ...) {
    char msg[400];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof msg, fmt, ap);
    va_end(ap);
    snprintf(g_err, sizeof g_err, "%s() error: %s", fn, msg);
    fprintf(stderr, "%s\n", g_err);  // same shape as the reference's diagnostics (tree.c:181)
}

static int do_init(int device, const char* fn) {
    if (g_ctx) return SFE_SUCCESS;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(fn, "no CUDA device available (%s); the engine has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return SFE_FAILURE;
    }
    if (device < 0) {
        const char* env = getenv("SFE_DEVICE");
        if (!env) env = getenv("LOCAL_RANK");
        device = env ? atoi(env) : 0;
    }
    if (device >= count) { set_error(fn, "device %d requested but only %d visible", device, count); return SFE_FAILURE; }
    cudaDeviceProp prop;
    SFE_CUDA(fn, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(fn, "device %d is sm_%d%d; this engine is built for sm_100a only", device, prop.major, prop.minor);
        return SFE_FAILURE;
    }
    SFE_CUDA(fn, cudaSetDevice(device));
    Ctx* c = new Ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    c->clock_khz = khz;
    SFE_CUDA(fn, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {   // the side streams carry the moment folds: short kernels that must not queue behind the long ones of the engine stream
        int lo = 0, hi = 0;
        SFE_CUDA(fn, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (auto& sd : c->side) SFE_CUDA(fn, cudaStreamCreateWithPriority(&sd, cudaStreamNonBlocking, hi));
        for (auto& sd : c->side2) SFE_CUDA(fn, cudaStreamCreateWithPriority(&sd, cudaStreamNonBlocking, hi));
    }
    for (auto& ev : c->ev) SFE_CUDA(fn, cudaEventCreate(&ev));
    memset(&c->timings, 0, sizeof c->timings);
    g_ctx = c;
    return SFE_SUCCESS;
}

int ensure_init(const char* fn) { return g_ctx ? SFE_SUCCESS : do_init(-1, fn); }

// ---------------------------------------------------------------------------- staging
static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static int ensure_bounce(const char* fn, size_t bytes) {
    Ctx* c = g_ctx;
    if (bytes <= c->bounce_cap) return SFE_SUCCESS;
    if (c->bounce) cudaFreeHost(c->bounce);
    c->bounce = nullptr; c->bounce_cap = 0;
    SFE_CUDA(fn, cudaMallocHost(&c->bounce, bytes));
    c->bounce_cap = bytes;
    return SFE_SUCCESS;
}

int h2d(const char* fn, void* dst, const void* src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return SFE_SUCCESS;
    g_ctx->timings.h2d_bytes += bytes;
    if (is_pinned(src)) {
        SFE_CUDA(fn, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
        return SFE_SUCCESS;
    }
    // pageable memory: stage through the pinned bounce buffer in 16 MiB pieces
    const size_t piece = (size_t)16 << 20;
    if (ensure_bounce(fn, 2 * piece)) return SFE_FAILURE;
    size_t off = 0; int half = 0;
    while (off < bytes) {
        size_t n = bytes - off < piece ? bytes - off : piece;
        char* b = (char*)g_ctx->bounce + half * piece;
        // the previous copy out of this half must be done before it is overwritten
        if (off >= 2 * piece) SFE_CUDA(fn, cudaStreamSynchronize(s));
        memcpy(b, (const char*)src + off, n);
        SFE_CUDA(fn, cudaMemcpyAsync((char*)dst + off, b, n, cudaMemcpyHostToDevice, s));
        off += n; half ^= 1;
    }
    SFE_CUDA(fn, cudaStreamSynchronize(s));
    return SFE_SUCCESS;
}

int d2h(const char* fn, void* dst, const void* src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return SFE_SUCCESS;
    g_ctx->timings.d2h_bytes += bytes;
    if (is_pinned(dst)) {
        SFE_CUDA(fn, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
        return SFE_SUCCESS;
    }
    const size_t piece = (size_t)16 << 20;
    if (ensure_bounce(fn, 2 * piece)) return SFE_FAILURE;
    size_t off = 0;
    while (off < bytes) {
        size_t n = bytes - off < piece ? bytes - off : piece;
        SFE_CUDA(fn, cudaMemcpyAsync(g_ctx->bounce, (const char*)src + off, n, cudaMemcpyDeviceToHost, s));
        SFE_CUDA(fn, cudaStreamSynchronize(s));
        memcpy((char*)dst + off, g_ctx->bounce, n);
        off += n;
    }
    return SFE_SUCCESS;
}

// event helpers for sfe_timings_t: ev[0]=start, ev[1]=after h2d, ev[2]=after build, ev[3]=after solve, ev[4]=end
static void t_begin(Ctx* c) {
    memset(&c->timings, 0, sizeof c->timings);
    c->launches = 0;
    cudaEventRecord(c->ev[0], c->stream);
}
static void t_mark(Ctx* c, int k) { cudaEventRecord(c->ev[k], c->stream); }
static int t_end(Ctx* c, const char* fn, bool has_build) {
    cudaEventRecord(c->ev[4], c->stream);
    SFE_CUDA(fn, cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&c->timings.h2d_ms, c->ev[0], c->ev[1]);
    if (has_build) {
        cudaEventElapsedTime(&c->timings.build_ms, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&c->timings.solve_ms, c->ev[2], c->ev[3]);
    } else {
        cudaEventElapsedTime(&c->timings.solve_ms, c->ev[1], c->ev[3]);
    }
    cudaEventElapsedTime(&c->timings.d2h_ms, c->ev[3], c->ev[4]);
    cudaEventElapsedTime(&c->timings.total_ms, c->ev[0], c->ev[4]);
    c->timings.launches = c->launches;
    return SFE_SUCCESS;
}

// `for_tree`: the upload feeds sfe_tree_build().  Any other upload overwrites (or reallocates) the records the device
// tree was built from, so the tree of an earlier sfe_tree_build() is retired: sfe_bh() / sfe_bh_overlay() then fail
// instead of walking a stale tree against different particles.
static int upload_particles(Ctx* c, const char* fn, const float* pos, const float* mas, uint32_t N, bool for_tree = false) {
    if (!for_tree && c->tree.valid && (c->tree.pm4_of_build == c->d_pm4.p || c->tree.pm4_of_build == nullptr)) {
        c->tree.valid = false; c->tree.pm4_of_build = nullptr;
    }
    if (c->d_pos.reserve((size_t)N * 12, fn) || c->d_mas.reserve((size_t)N * 4, fn) ||
        c->d_pm4.reserve((size_t)N * 16, fn))
        return SFE_FAILURE;
    if (h2d(fn, c->d_pos.p, pos, (size_t)N * 12, c->stream)) return SFE_FAILURE;
    if (h2d(fn, c->d_mas.p, mas, (size_t)N * 4, c->stream)) return SFE_FAILURE;
    return launch_pack_pm4(c->d_pos.as<float>(), c->d_mas.as<float>(), c->d_pm4.as<float4>(), N, c->stream);
}

}  // namespace sfe

using namespace sfe;

// ============================================================================ lifecycle
extern "C" int sfe_init(int device) { return do_init(device, "sfe_init"); }

extern "C" int sfe_shutdown(void) {
    Ctx* c = g_ctx;
    if (!c) return SFE_SUCCESS;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto& kv : g_registered) cudaHostUnregister(kv.first);
    g_registered.clear();
    saver_release();
    sim_release(c);
    peer_release(c);
    Tree& t = c->tree;
    DevBuf* bufs[] = {&c->d_pos, &c->d_mas, &c->d_pm4, &c->d_acc, &c->d_pot, &c->d_rad, &c->d_rho, &c->d_prs,
                      &c->pp_src, &c->pp_partial, &c->pp_flag, &c->sph_src, &c->sph_partial,
                      &c->sph_cells.desc, &c->sph_cells.count, &c->sph_cells.start, &c->sph_cells.key[0], &c->sph_cells.key[1],
                      &c->sph_cells.val[0], &c->sph_cells.val[1], &c->sph_cells.hist, &c->sph_cells.rec, &c->sph_cells.tlist, &c->sph_cells.part, &c->sph_cells.arrive,
                      &t.nodes, &t.seg, &t.next, &t.child_base, &t.counts, &t.keys, &t.perm_out, &t.oct, &t.rank, &t.nmid, &t.lookback, &t.xl_sum, &t.xl_map, &t.xl_flag, &t.xl_list, &t.scan_tmp, &t.slice_count, &t.root_gate, &t.meta, &t.rec0, &t.rec1, &t.rec2, &t.recp, &t.bfs_of, &t.recp_par, &t.recp_hist, &t.recp_pref,
                      &t.sub_list, &t.bounds_buf, &t.top_table, &t.top_cols, &t.top_rank};
    for (DevBuf& b : t.srec) b.release();
    if (t.meta_host) { cudaFreeHost(t.meta_host); t.meta_host = nullptr; }
    if (t.summary_host) { cudaFreeHost(t.summary_host); t.summary_host = nullptr; t.summary_dev = nullptr; }
    if (t.ev_init) {
        for (auto& ev : t.ev_perm) cudaEventDestroy(ev);
        for (auto& ev : t.ev_side) cudaEventDestroy(ev);
        for (auto& ev : t.ev_pre) cudaEventDestroy(ev);
        for (auto& ev : t.ev_pre2) cudaEventDestroy(ev);
        t.ev_init = false;
    }
    t.valid = false;
    for (DevBuf* b : bufs) b->release();
    c->ov_lines.release(); c->ov_rows.release();
    c->vis_atomic.release(); c->vis_render.release(); c->vis_info.release(); c->vis_pm4.release();
    if (c->ov_info) cudaFreeHost(c->ov_info);
    if (c->bounce) cudaFreeHost(c->bounce);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    for (auto& sd : c->side) if (sd) cudaStreamDestroy(sd);
    for (auto& sd : c->side2) if (sd) cudaStreamDestroy(sd);
    delete c;
    g_ctx = nullptr;
    return SFE_SUCCESS;
}

extern "C" const char* sfe_last_error(void) { return g_err; }

extern "C" int sfe_host_register(void* ptr, uint64_t bytes) {
    const char* fn = "sfe_host_register";
    if (ensure_init(fn)) return SFE_FAILURE;
    // ranges may nest or overlap (a shim that first sees pos / mas and later the whole sim.mem block): a range inside an
    // existing registration needs nothing; existing registrations the new range touches are given up first
    const char* lo = (const char*)ptr; const char* hi = lo + bytes;
    for (auto it = g_registered.begin(); it != g_registered.end();) {
        const char* a = (const char*)it->first; const char* b = a + it->second;
        if (lo >= a && hi <= b) return SFE_SUCCESS;
        if (a < hi && lo < b) { cudaHostUnregister(it->first); it = g_registered.erase(it); } else ++it;
    }
    if (is_pinned(ptr) && is_pinned(hi - 1)) return SFE_SUCCESS;  // already pinned by the caller
    SFE_CUDA(fn, cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    g_registered[ptr] = bytes;
    return SFE_SUCCESS;
}

extern "C" int sfe_host_unregister(void* ptr) {
    auto it = g_registered.find(ptr);
    if (it == g_registered.end()) return SFE_SUCCESS;
    cudaHostUnregister(ptr);
    g_registered.erase(it);
    return SFE_SUCCESS;
}

extern "C" int sfe_get_timings(sfe_timings_t* out) {
    if (!g_ctx || !out) return SFE_FAILURE;
    *out = g_ctx->timings;
    return SFE_SUCCESS;
}

extern "C" int sfe_device_info(int* sm_count, int* clock_khz, double* fp32_peak_tflops) {
    if (ensure_init("sfe_device_info")) return SFE_FAILURE;
    if (sm_count) *sm_count = g_ctx->sm_count;
    if (clock_khz) *clock_khz = g_ctx->clock_khz;
    if (fp32_peak_tflops) *fp32_peak_tflops = 2.0 * 128.0 * g_ctx->sm_count * (double)g_ctx->clock_khz * 1e3 / 1e12;
    return SFE_SUCCESS;
}

extern "C" uint64_t sfe_launch_count(void) { return g_ctx ? g_ctx->total_launches : 0; }

// ============================================================================ host entry points
extern "C" int sfe_pp_slice(const float* pos, const float* mas, float* acc, float* pot, uint32_t N,
                            uint32_t i0, uint32_t ni, float G, float eps2) {
    const char* fn = "sfe_pp";
    if (ensure_init(fn)) return SFE_FAILURE;
    if (!pos || !mas || !acc || !pot) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if ((uint64_t)i0 + ni > N) { set_error(fn, "target slice [%u,%u) exceeds N=%u", i0, i0 + ni, N); return SFE_FAILURE; }
    if (N == 0 || ni == 0) return SFE_SUCCESS;
    Ctx* c = g_ctx;
    t_begin(c);
    if (upload_particles(c, fn, pos, mas, N)) return SFE_FAILURE;
    if (c->d_acc.reserve((size_t)N * 12, fn) || c->d_pot.reserve((size_t)N * 4, fn)) return SFE_FAILURE;
    t_mark(c, 1);
    if (launch_pp(c, c->d_pm4.as<float4>(), N, i0, ni, c->d_acc.as<float>(), c->d_pot.as<float>(), G, eps2, c->stream))
        return SFE_FAILURE;
    t_mark(c, 3);
    if (d2h(fn, acc + 3u * (size_t)i0, c->d_acc.as<float>() + 3u * (size_t)i0, (size_t)ni * 12, c->stream)) return SFE_FAILURE;
    if (d2h(fn, pot + i0, c->d_pot.as<float>() + i0, (size_t)ni * 4, c->stream)) return SFE_FAILURE;
    return t_end(c, fn, false);
}

extern "C" int sfe_pp(const float* pos, const float* mas, float* acc, float* pot, uint32_t N, float G, float eps2) {
    return sfe_pp_slice(pos, mas, acc, pot, N, 0, N, G, eps2);
}

extern "C" int sfe_tree_build(const float bounds[6], const float* pos, const float* mas, uint32_t N,
                              int depth_max, uint32_t nodes_max) {
    const char* fn = "sfe_tree_build";
    if (ensure_init(fn)) return SFE_FAILURE;
    if (!bounds || ((!pos || !mas) && N > 0)) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    Ctx* c = g_ctx;
    t_begin(c);
    c->tree.valid = false;
    if (upload_particles(c, fn, pos, mas, N, /*for_tree=*/true)) return SFE_FAILURE;
    t_mark(c, 1);
    if (launch_tree_build(c, bounds, c->d_pm4.as<float4>(), N, depth_max, nodes_max, c->stream)) return SFE_FAILURE;
    t_mark(c, 2);
    t_mark(c, 3);
    if (t_end(c, fn, true)) return SFE_FAILURE;
    return tree_status(c, c->stream);
}

extern "C" int sfe_bh_slice(const float* pos, const float* mas, float* acc, float* pot, uint32_t N,
                            uint32_t i0, uint32_t ni, float G, float eps2, float theta, uint32_t flavour) {
    const char* fn = "sfe_bh";
    if (ensure_init(fn)) return SFE_FAILURE;
    if (!pos || !mas || !acc || !pot) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if ((uint64_t)i0 + ni > N) { set_error(fn, "target slice [%u,%u) exceeds N=%u", i0, i0 + ni, N); return SFE_FAILURE; }
    Ctx* c = g_ctx;
    if (!c->tree.valid || c->tree.N != N) {
        set_error(fn, "no tree for N=%u: call sfe_tree_build() on the same particles first", N);
        return SFE_FAILURE;
    }
    if (N == 0 || ni == 0) return SFE_SUCCESS;
    t_begin(c);
    // the particles were uploaded by sfe_tree_build() (the reference calls build and walk back to back on
    // the same arrays, simulation.c:312-321); they are not copied again.
    if (c->d_acc.reserve((size_t)N * 12, fn) || c->d_pot.reserve((size_t)N * 4, fn)) return SFE_FAILURE;
    t_mark(c, 1);
    if (launch_bh(c, c->d_pm4.as<float4>(), N, i0, ni, c->d_acc.as<float>(), c->d_pot.as<float>(), G, eps2, theta,
                  flavour, c->stream))
        return SFE_FAILURE;
    t_mark(c, 3);
    if (d2h(fn, acc + 3u * (size_t)i0, c->d_acc.as<float>() + 3u * (size_t)i0, (size_t)ni * 12, c->stream)) return SFE_FAILURE;
    if (d2h(fn, pot + i0, c->d_pot.as<float>() + i0, (size_t)ni * 4, c->stream)) return SFE_FAILURE;
    return t_end(c, fn, false);
}

extern "C" int sfe_bh(const float* pos, const float* mas, float* acc, float* pot, uint32_t N,
                      float G, float eps2, float theta, uint32_t flavour) {
    return sfe_bh_slice(pos, mas, acc, pot, N, 0, N, G, eps2, theta, flavour);
}

// The SOLVER_TREE branch of sim_solv() (simulation.c:245-325) in one call on host arrays: root box, tree_build,
// solve_gravity_part_tree.  Nothing waits for the host between the upload and the download but the root box (six floats
// through mapped memory, because the box is a kernel argument of the build); the build's status is raised at the end
// (a walk on a failed build returns at once on the device).
extern "C" int sfe_solv_tree(const float* pos, const float* mas, float* acc, float* pot, uint32_t N, uint32_t step_number,
                             int depth_max, uint32_t nodes_max, float G, float eps2, float theta, uint32_t flavour,
                             float bounds_out[6]) {
    const char* fn = "sfe_solv_tree";
    if (ensure_init(fn)) return SFE_FAILURE;
    if (!pos || !mas || !acc || !pot) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if (N == 0) { set_error(fn, "N must be at least 1 (the reference divides by N, simulation.c:277)"); return SFE_FAILURE; }
    Ctx* c = g_ctx;
    t_begin(c);
    c->tree.valid = false;
    if (upload_particles(c, fn, pos, mas, N, /*for_tree=*/true)) return SFE_FAILURE;
    if (c->d_acc.reserve((size_t)N * 12, fn) || c->d_pot.reserve((size_t)N * 4, fn)) return SFE_FAILURE;
    t_mark(c, 1);
    float b[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* box_dev = nullptr;                      // the box stays on the device; the host copy is picked up at the end
    if (launch_root_box_dev(c, c->d_pm4.as<float4>(), N, step_number, &box_dev, c->stream)) return SFE_FAILURE;
    if (launch_tree_build(c, b, c->d_pm4.as<float4>(), N, depth_max, nodes_max, c->stream, box_dev)) return SFE_FAILURE;
    t_mark(c, 2);
    if (launch_bh(c, c->d_pm4.as<float4>(), N, 0, N, c->d_acc.as<float>(), c->d_pot.as<float>(), G, eps2, theta, flavour, c->stream))
        return SFE_FAILURE;
    t_mark(c, 3);
    if (d2h(fn, acc, c->d_acc.p, (size_t)N * 12, c->stream)) return SFE_FAILURE;
    if (d2h(fn, pot, c->d_pot.p, (size_t)N * 4, c->stream)) return SFE_FAILURE;
    if (t_end(c, fn, true)) return SFE_FAILURE;
    if (bounds_out && root_box_host(c, bounds_out, c->stream, fn)) return SFE_FAILURE;
    return tree_status(c, c->stream);
}

extern "C" int sfe_sph(const float* pos, const float* mas, const float* rad, float* rho, float* prs, float* acc,
                       uint32_t N, float eos_k, float eos_n, float cutoff2) {
    const char* fn = "sfe_sph";
    if (ensure_init(fn)) return SFE_FAILURE;
    if (!pos || !mas || !rad || !rho || !prs || !acc) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if (N == 0) return SFE_SUCCESS;
    Ctx* c = g_ctx;
    t_begin(c);
    if (upload_particles(c, fn, pos, mas, N)) return SFE_FAILURE;
    if (c->d_rad.reserve((size_t)N * 4, fn) || c->d_rho.reserve((size_t)N * 4, fn) ||
        c->d_prs.reserve((size_t)N * 4, fn) || c->d_acc.reserve((size_t)N * 12, fn))
        return SFE_FAILURE;
    if (h2d(fn, c->d_rad.p, rad, (size_t)N * 4, c->stream)) return SFE_FAILURE;
    if (h2d(fn, c->d_acc.p, acc, (size_t)N * 12, c->stream)) return SFE_FAILURE;  // acc is read-modify-write (sph.c:274-276)
    t_mark(c, 1);
    float4* pm4 = c->d_pm4.as<float4>();
    if (launch_sph_density(c, pm4, c->d_rad.as<float>(), N, 0, N, c->d_rho.as<float>(), c->stream)) return SFE_FAILURE;
    if (launch_sph_eos(c->d_rho.as<float>(), c->d_prs.as<float>(), N, eos_k, eos_n, c->stream)) return SFE_FAILURE;
    if (launch_sph_force(c, pm4, c->d_rad.as<float>(), c->d_rho.as<float>(), c->d_prs.as<float>(), N, 0, N, cutoff2,
                         c->d_acc.as<float>(), c->stream, /*reuse_cells=*/true))   // same snapshot as the density pass
        return SFE_FAILURE;
    t_mark(c, 3);
    if (d2h(fn, rho, c->d_rho.p, (size_t)N * 4, c->stream)) return SFE_FAILURE;
    if (d2h(fn, prs, c->d_prs.p, (size_t)N * 4, c->stream)) return SFE_FAILURE;
    if (d2h(fn, acc, c->d_acc.p, (size_t)N * 12, c->stream)) return SFE_FAILURE;
    return t_end(c, fn, false);
}

// ============================================================================ device entry points
extern "C" int sfe_dev_pack_pm4(const float* pos, const float* mas, void* pm4, uint32_t n, void* stream) {
    if (ensure_init("sfe_dev_pack_pm4")) return SFE_FAILURE;
    return launch_pack_pm4(pos, mas, (float4*)pm4, n, (cudaStream_t)stream);
}

extern "C" int sfe_dev_pp(const void* pm4, uint32_t N, uint32_t i0, uint32_t ni, float* acc, float* pot,
                          float G, float eps2, void* stream) {
    if (ensure_init("sfe_dev_pp")) return SFE_FAILURE;
    return launch_pp(g_ctx, (const float4*)pm4, N, i0, ni, acc, pot, G, eps2, (cudaStream_t)stream);
}

// root box of sim_solv() (simulation.c:245-301) + tree_build in one call, the box never leaving the device
extern "C" int sfe_dev_tree_build_box(const void* pm4, uint32_t N, uint32_t step_number, int depth_max, uint32_t nodes_max, void* stream) {
    const char* fn = "sfe_dev_tree_build_box";
    if (ensure_init(fn)) return SFE_FAILURE;
    if (!pm4) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    const float* box_dev = nullptr;
    const float unused[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (launch_root_box_dev(g_ctx, (const float4*)pm4, N, step_number, &box_dev, (cudaStream_t)stream)) return SFE_FAILURE;
    return launch_tree_build(g_ctx, unused, (const float4*)pm4, N, depth_max, nodes_max, (cudaStream_t)stream, box_dev);
}
// the box of the last sfe_dev_tree_build_box() / sfe_dev_root_box(), once the stream has computed it (waits for it)
extern "C" int sfe_dev_last_box(float bounds[6], void* stream) {
    const char* fn = "sfe_dev_last_box";
    if (ensure_init(fn)) return SFE_FAILURE;
    if (!bounds) { set_error(fn, "null pointer argument"); return SFE_FAILURE; }
    if (!g_ctx->sim.box_host || g_ctx->sim.box_seq == 0u) { set_error(fn, "no root box has been computed"); return SFE_FAILURE; }
    return root_box_host(g_ctx, bounds, (cudaStream_t)stream, fn);
}

extern "C" int sfe_dev_tree_build(const float bounds[6], const void* pm4, uint32_t N, int depth_max,
                                  uint32_t nodes_max, void* stream) {
    if (ensure_init("sfe_dev_tree_build")) return SFE_FAILURE;
    return launch_tree_build(g_ctx, bounds, (const float4*)pm4, N, depth_max, nodes_max, (cudaStream_t)stream);
}

extern "C" int sfe_dev_tree_status(void* stream, uint32_t* num_nodes, uint32_t* num_inbox, uint32_t* depth) {
    if (ensure_init("sfe_dev_tree_status")) return SFE_FAILURE;
    int rc = tree_status(g_ctx, (cudaStream_t)stream);
    if (num_nodes) *num_nodes = g_ctx->tree.num_nodes;
    if (num_inbox) *num_inbox = g_ctx->tree.num_inbox;
    if (depth) *depth = g_ctx->tree.depth;
    return rc;
}

extern "C" int sfe_dev_bh(const void* pm4, uint32_t N, uint32_t i0, uint32_t ni, float* acc, float* pot,
                          float G, float eps2, float theta, uint32_t flavour, void* stream) {
    if (ensure_init("sfe_dev_bh")) return SFE_FAILURE;
    return launch_bh(g_ctx, (const float4*)pm4, N, i0, ni, acc, pot, G, eps2, theta, flavour, (cudaStream_t)stream);
}

extern "C" int sfe_dev_sph_density(const void* pm4, const float* rad, uint32_t N, uint32_t i0, uint32_t ni,
                                   float* rho, void* stream) {
    if (ensure_init("sfe_dev_sph_density")) return SFE_FAILURE;
    return launch_sph_density(g_ctx, (const float4*)pm4, rad, N, i0, ni, rho, (cudaStream_t)stream);
}

extern "C" int sfe_sph_pairs_evaluated(uint64_t* pairs, void* stream) {
    if (ensure_init("sfe_sph_pairs_evaluated")) return SFE_FAILURE;
    if (!pairs) { set_error("sfe_sph_pairs_evaluated", "null pointer argument"); return SFE_FAILURE; }
    unsigned long long v = 0ull;
    // (`stream` as given: NULL is the legacy default stream, which is what a caller on that stream has enqueued the passes on)
    if (cells_stats(g_ctx, &v, (cudaStream_t)stream, "sfe_sph_pairs_evaluated")) return SFE_FAILURE;
    *pairs = v;
    return SFE_SUCCESS;
}

extern "C" int sfe_dev_sph_eos(const float* rho, float* prs, uint32_t n, float eos_k, float eos_n, void* stream) {
    if (ensure_init("sfe_dev_sph_eos")) return SFE_FAILURE;
    return launch_sph_eos(rho, prs, n, eos_k, eos_n, (cudaStream_t)stream);
}

extern "C" int sfe_dev_sph_force(const void* pm4, const float* rad, const float* rho, const float* prs, uint32_t N,
                                 uint32_t i0, uint32_t ni, float cutoff2, float* acc, void* stream) {
    if (ensure_init("sfe_dev_sph_force")) return SFE_FAILURE;
    return launch_sph_force(g_ctx, (const float4*)pm4, rad, rho, prs, N, i0, ni, cutoff2, acc, (cudaStream_t)stream);
}
