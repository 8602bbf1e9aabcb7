// Host matrix -> device memory for PAGEABLE sources (the end-to-end figure).  The reference hands the
// grouping stage an ordinary numpy array (main.py:71); cudaMemcpy from pageable memory stages through the driver's own bounce
// buffer with one host thread and reaches 6-13 GB/s where the link carries 55.  Here n native threads copy 8 MB pieces into
// their own page-locked double buffers and queue the piece's cudaMemcpyAsync on their own streams: the host copies run in
// parallel and overlap the DMA.  The caller's stream waits for all pieces; nothing is ordered before that.
// Host code only (no kernels); pinned sources do not come here (one plain async copy is already at link speed).
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include "common.cuh"

namespace vtf {

constexpr size_t kUpPiece = 8u << 20;
constexpr int kUpMaxThreads = 16;

struct UploadCtx {
    int dev = -1;
    int n = 0;
    void* pin[2 * kUpMaxThreads] = {};
    cudaStream_t st[kUpMaxThreads] = {};
    cudaEvent_t ev[2 * kUpMaxThreads] = {};
    bool used[2 * kUpMaxThreads] = {};
    cudaEvent_t start = nullptr;
};
static UploadCtx g_up;
static std::mutex g_up_mutex;

static int upload_ctx(int n_threads) {
    int dev = 0;
    VTF_CUDA(cudaGetDevice(&dev), "vtf_upload_pageable");
    if (g_up.dev >= 0 && g_up.dev != dev) return fail(VTF_ERR_ARG, "vtf_upload_pageable", "one device per process");
    g_up.dev = dev;
    if (!g_up.start) VTF_CUDA(cudaEventCreateWithFlags(&g_up.start, cudaEventDisableTiming), "vtf_upload_pageable");
    for (; g_up.n < n_threads; ++g_up.n) {
        const int w = g_up.n;
        VTF_CUDA(cudaStreamCreateWithFlags(&g_up.st[w], cudaStreamNonBlocking), "vtf_upload_pageable/stream");
        for (int s = 0; s < 2; ++s) {
            VTF_CUDA(cudaHostAlloc(&g_up.pin[2 * w + s], kUpPiece, cudaHostAllocDefault), "vtf_upload_pageable/pinned");
            VTF_CUDA(cudaEventCreateWithFlags(&g_up.ev[2 * w + s], cudaEventDisableTiming), "vtf_upload_pageable/event");
        }
    }
    return VTF_OK;
}

}  // namespace vtf

using namespace vtf;

extern "C" int vtf_upload_pageable(const void* src_host, void* dst_dev, size_t bytes, int n_threads, void* stream) {
    VTF_REQUIRE(src_host && dst_dev && n_threads > 0, "vtf_upload_pageable");
    if (bytes == 0) return VTF_OK;
    std::lock_guard<std::mutex> lock(g_up_mutex);
    const size_t n_pieces = (bytes + kUpPiece - 1) / kUpPiece;
    const int T = (int)std::min<size_t>((size_t)std::min(n_threads, kUpMaxThreads), n_pieces);
    int rc = upload_ctx(T);
    if (rc != VTF_OK) return rc;
    // the copies may start once the caller's stream has reached this point (dst may have been in use before)
    VTF_CUDA(cudaEventRecord(g_up.start, (cudaStream_t)stream), "vtf_upload_pageable");
    int errs[kUpMaxThreads] = {};
    const int dev = g_up.dev;
    auto worker = [&](int w) {
        if (cudaSetDevice(dev) != cudaSuccess || cudaStreamWaitEvent(g_up.st[w], g_up.start, 0) != cudaSuccess) { errs[w] = 1; return; }
        int k = 0;
        for (size_t c = (size_t)w; c < n_pieces; c += (size_t)T, ++k) {
            const int slot = 2 * w + (k & 1);
            // the slot's previous piece (of this call or of an earlier one) must have left the buffer
            if (g_up.used[slot] && cudaEventSynchronize(g_up.ev[slot]) != cudaSuccess) { errs[w] = 1; return; }
            const size_t off = c * kUpPiece, n = std::min(kUpPiece, bytes - off);
            memcpy(g_up.pin[slot], (const char*)src_host + off, n);
            if (cudaMemcpyAsync((char*)dst_dev + off, g_up.pin[slot], n, cudaMemcpyHostToDevice, g_up.st[w]) != cudaSuccess ||
                cudaEventRecord(g_up.ev[slot], g_up.st[w]) != cudaSuccess) { errs[w] = 1; return; }
            g_up.used[slot] = true;
        }
    };
    std::vector<std::thread> pool;
    for (int w = 1; w < T; ++w) pool.emplace_back(worker, w);
    worker(0);
    for (auto& t : pool) t.join();
    for (int w = 0; w < T; ++w)
        if (errs[w]) { cudaGetLastError(); return fail(VTF_ERR_CUDA, "vtf_upload_pageable", "a piece failed to copy"); }
    // the caller's stream continues after the last piece of every worker
    for (int w = 0; w < T; ++w) {
        const size_t mine = (n_pieces - (size_t)w + (size_t)T - 1) / (size_t)T;
        if (mine == 0) continue;
        const int last = 2 * w + (int)((mine - 1) & 1);
        VTF_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, g_up.ev[last], 0), "vtf_upload_pageable/join");
    }
    return VTF_OK;
}
