// sfe_save.cu - RAW snapshot writer fed from the resident state (sim_save_raw, simulation_io.c:81-149).
//
// The reference's snapshot is the state block itself: real[N * SIM_DOF], block p at element N*p (simulation.c:52),
// written with one fwrite().  With the state resident in HBM the file is produced without stopping the leapfrog:
//
//   compute stream   the rank's rows of every block are copied into a device-side snapshot slot (pos/mas unpacked
//                    from the {x,y,z,m} records): a few hundred microseconds of device copies, after which the next
//                    step may overwrite the state;
//   copy stream      one device-to-host copy of the slot into pinned memory, behind an event;
//   writer thread    waits for that event and pwrite()s each block's rows at its offset in the file.
//
// Two slots alternate, so a snapshot can be in flight while the next one is taken; a third request waits for the
// older one.  In multi-rank mode (sfe_peer.cu) every rank holds the per-particle blocks of its own target slice only,
// and every rank writes exactly those rows: the ranks of one node fill one file in parallel, nothing is gathered.
#include "sfe_internal.cuh"
#include <condition_variable>
#include <deque>
#include <fcntl.h>
#include <mutex>
#include <string>
#include <thread>
#include <errno.h>
#include <unistd.h>

namespace sfe {

struct SaveSeg { uint64_t file_off, bytes; };
struct SaveJob {
    std::string path;
    int slot;
    uint64_t file_bytes;
    SaveSeg seg[10];
    int nseg;
};

struct Saver {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::deque<SaveJob> q;
    bool started = false, stop = false;
    bool busy[2] = {false, false};
    int failed = 0;
    char errmsg[400] = "";
    int device = 0;
    cudaStream_t copy = nullptr;
    cudaEvent_t ev_snap[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    void* dev[2] = {nullptr, nullptr};
    void* host[2] = {nullptr, nullptr};
    size_t cap = 0;
    int next_slot = 0;
};
// Heap-allocated and never destroyed: a process that exits without sfe_shutdown() must not run the destructors of a
// condition variable / thread that the (then still waiting) writer uses.
static Saver& g_saver = *new Saver();

static void saver_fail(Saver& sv, const char* what, const char* path, int err) {
    std::lock_guard<std::mutex> lk(sv.m);
    if (!sv.failed) snprintf(sv.errmsg, sizeof sv.errmsg, "%s(\"%s\") failed: %s", what, path, strerror(err));
    sv.failed = 1;
}

static void saver_main(Saver* svp) {
    Saver& sv = *svp;
    cudaSetDevice(sv.device);
    for (;;) {
        SaveJob job;
        {
            std::unique_lock<std::mutex> lk(sv.m);
            sv.cv.wait(lk, [&] { return sv.stop || !sv.q.empty(); });
            if (sv.q.empty()) return;
            job = sv.q.front();
        }
        cudaError_t e = cudaEventSynchronize(sv.ev_done[job.slot]);
        if (e != cudaSuccess) {
            std::lock_guard<std::mutex> lk(sv.m);
            if (!sv.failed) snprintf(sv.errmsg, sizeof sv.errmsg, "device copy for \"%s\" failed: %s", job.path.c_str(), cudaGetErrorString(e));
            sv.failed = 1;
        } else {
            // no O_TRUNC: in multi-rank mode the other ranks write their rows into the same file
            int fd = open(job.path.c_str(), O_WRONLY | O_CREAT, 0644);
            if (fd < 0) saver_fail(sv, "open", job.path.c_str(), errno);
            else {
                if (ftruncate(fd, (off_t)job.file_bytes) != 0) saver_fail(sv, "ftruncate", job.path.c_str(), errno);
                const char* src = (const char*)sv.host[job.slot];
                for (int k = 0; k < job.nseg; k++) {
                    uint64_t done = 0;
                    while (done < job.seg[k].bytes) {
                        ssize_t w = pwrite(fd, src + done, job.seg[k].bytes - done, (off_t)(job.seg[k].file_off + done));
                        if (w < 0) { if (errno == EINTR) continue; saver_fail(sv, "pwrite", job.path.c_str(), errno); break; }
                        done += (uint64_t)w;
                    }
                    src += job.seg[k].bytes;
                }
                if (close(fd) != 0) saver_fail(sv, "close", job.path.c_str(), errno);
            }
        }
        {
            std::lock_guard<std::mutex> lk(sv.m);
            sv.q.pop_front();
            sv.busy[job.slot] = false;
        }
        sv.cv.notify_all();
    }
}

static int saver_drain(Saver& sv, const char* fn) {
    if (!sv.started) return SFE_SUCCESS;
    std::unique_lock<std::mutex> lk(sv.m);
    sv.cv.wait(lk, [&] { return sv.q.empty(); });
    if (sv.failed) {
        set_error(fn, "%s", sv.errmsg);
        sv.failed = 0;
        return SFE_FAILURE;
    }
    return SFE_SUCCESS;
}

void saver_release() {
    Saver& sv = g_saver;
    if (sv.started) {
        { std::unique_lock<std::mutex> lk(sv.m); sv.cv.wait(lk, [&] { return sv.q.empty(); }); sv.stop = true; }
        sv.cv.notify_all();
        sv.th.join();
        sv.started = false; sv.stop = false;
    }
    for (int k = 0; k < 2; k++) {
        if (sv.dev[k]) cudaFree(sv.dev[k]);
        if (sv.host[k]) cudaFreeHost(sv.host[k]);
        if (sv.ev_snap[k]) cudaEventDestroy(sv.ev_snap[k]);
        if (sv.ev_done[k]) cudaEventDestroy(sv.ev_done[k]);
        sv.dev[k] = sv.host[k] = nullptr; sv.ev_snap[k] = sv.ev_done[k] = nullptr; sv.busy[k] = false;
    }
    if (sv.copy) cudaStreamDestroy(sv.copy);
    sv.copy = nullptr; sv.cap = 0; sv.failed = 0;
}

// rows [i0, i0+ni) of the {x,y,z,m} records -> xyz-interleaved positions and masses
__global__ void save_unpack_kernel(const float4* __restrict__ pm4, uint32_t i0, uint32_t ni, float* __restrict__ pos,
                                   float* __restrict__ mas) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ni) return;
    const float4 p = pm4[i0 + t];
    pos[3u * t] = p.x; pos[3u * t + 1u] = p.y; pos[3u * t + 2u] = p.z;
    mas[t] = p.w;
}

}  // namespace sfe

using namespace sfe;

extern "C" int sfe_sim_save_raw_async(const char* filename) {
    const char* fn = "sfe_sim_save_raw_async";
    if (ensure_init(fn)) return SFE_FAILURE;
    Ctx* c = ctx();
    SimState& st = c->sim;
    if (!st.loaded) { set_error(fn, "no resident state: call sfe_sim_upload() first"); return SFE_FAILURE; }
    if (!filename) { set_error(fn, "filename is NULL!"); return SFE_FAILURE; }   // simulation_io.c:98-101
    Saver& sv = g_saver;
    const size_t ni = st.ni, n = st.N, a = st.i0;
    const size_t bytes = (size_t)st.dof * ni * sizeof(float);
    if (!sv.started) {
        sv.device = c->device;
        SFE_CUDA(fn, cudaStreamCreateWithFlags(&sv.copy, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            SFE_CUDA(fn, cudaEventCreateWithFlags(&sv.ev_snap[k], cudaEventDisableTiming));
            SFE_CUDA(fn, cudaEventCreateWithFlags(&sv.ev_done[k], cudaEventDisableTiming | cudaEventBlockingSync));
        }
        sv.th = std::thread(saver_main, &sv);
        sv.started = true;
    }
    if (bytes > sv.cap) {
        if (saver_drain(sv, fn)) return SFE_FAILURE;          // nothing may still read the old slots
        for (int k = 0; k < 2; k++) {
            if (sv.dev[k]) cudaFree(sv.dev[k]);
            if (sv.host[k]) cudaFreeHost(sv.host[k]);
            sv.dev[k] = sv.host[k] = nullptr;
        }
        sv.cap = 0;
        const size_t want = bytes + (bytes >> 3) + 256;
        for (int k = 0; k < 2; k++) {
            SFE_CUDA(fn, cudaMalloc(&sv.dev[k], want));
            SFE_CUDA(fn, cudaMallocHost(&sv.host[k], want));
        }
        sv.cap = want;
    }
    const int slot = sv.next_slot; sv.next_slot ^= 1;
    {   // the slot's previous snapshot must be on disk before it is overwritten
        std::unique_lock<std::mutex> lk(sv.m);
        sv.cv.wait(lk, [&] { return !sv.busy[slot]; });
        sv.busy[slot] = true;
    }
    SaveJob job;
    auto enqueue = [&]() -> int {
        cudaStream_t s = c->stream;
        float* d = (float*)sv.dev[slot];
        const float4* pm4 = st.peer_mode ? reinterpret_cast<const float4*>((const char*)c->peer.local + st.pm4_off[st.cur]) : st.pm4.as<float4>();
        const float* rho = st.peer_mode ? reinterpret_cast<const float*>((const char*)c->peer.local + st.rho_off) : st.rho.as<float>();
        const float* acc = st.peer_mode ? reinterpret_cast<const float*>((const char*)c->peer.local + st.acc_off) : st.acc.as<float>();
        const float* pot = st.peer_mode ? reinterpret_cast<const float*>((const char*)c->peer.local + st.pot_off) : st.pot.as<float>();
        // slot layout = the blocks in file order, each restricted to the rank's rows (enum sim_conf, simulation.h:17-51)
        struct Blk { uint32_t off, w; const float* src; };
        const Blk blk[10] = {{0, 3, nullptr}, {3, 3, st.vel.as<float>()}, {6, 3, acc}, {9, 1, nullptr},
                             {10, 1, pot}, {11, 1, st.ken.as<float>()}, {12, 1, st.pen.as<float>()},
                             {13, 1, st.rad.as<float>()}, {14, 1, rho}, {15, 1, st.prs.as<float>()}};
        const int nblk = st.dof == 16u ? 10 : 7;
        job.path = filename; job.slot = slot; job.nseg = nblk;
        job.file_bytes = (uint64_t)n * st.dof * sizeof(float);
        size_t at = 0, pos_at = 0, mas_at = 0;
        for (int k = 0; k < nblk; k++) {
            job.seg[k].file_off = ((uint64_t)n * blk[k].off + (uint64_t)blk[k].w * a) * sizeof(float);
            job.seg[k].bytes = (uint64_t)blk[k].w * ni * sizeof(float);
            if (blk[k].off == 0) pos_at = at;
            else if (blk[k].off == 9) mas_at = at;
            else if (ni) SFE_CUDA(fn, cudaMemcpyAsync(d + at, blk[k].src + (size_t)blk[k].w * a, job.seg[k].bytes, cudaMemcpyDeviceToDevice, s));
            at += (size_t)blk[k].w * ni;
        }
        if (ni) {
            save_unpack_kernel<<<(uint32_t)((ni + 255) / 256), 256, 0, s>>>(pm4, (uint32_t)a, (uint32_t)ni, d + pos_at, d + mas_at);
            if (check_launch(fn)) return SFE_FAILURE;
        }
        SFE_CUDA(fn, cudaEventRecord(sv.ev_snap[slot], s));
        SFE_CUDA(fn, cudaStreamWaitEvent(sv.copy, sv.ev_snap[slot], 0));
        if (bytes) SFE_CUDA(fn, cudaMemcpyAsync(sv.host[slot], sv.dev[slot], bytes, cudaMemcpyDeviceToHost, sv.copy));
        SFE_CUDA(fn, cudaEventRecord(sv.ev_done[slot], sv.copy));
        return SFE_SUCCESS;
    };
    if (enqueue()) {
        { std::lock_guard<std::mutex> lk(sv.m); sv.busy[slot] = false; }
        sv.cv.notify_all();
        return SFE_FAILURE;
    }
    {
        std::lock_guard<std::mutex> lk(sv.m);
        sv.q.push_back(job);
    }
    sv.cv.notify_all();
    return SFE_SUCCESS;
}

extern "C" int sfe_sim_save_wait(void) {
    if (!ctx()) return SFE_SUCCESS;
    return saver_drain(g_saver, "sfe_sim_save_wait");
}
