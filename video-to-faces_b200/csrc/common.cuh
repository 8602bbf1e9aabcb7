// Shared helpers for the vtf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/vtf_b200.h"

namespace vtf {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* what, const char* detail) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, detail ? detail : "");
    return code;
}

// Every kernel launch goes through this so that bench.py can report gpu_launches.
#define VTF_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                     \
        kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);            \
        vtf::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
    } while (0)

#define VTF_CHECK_LAST(what)                                                                 \
    do {                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess) return vtf::fail(VTF_ERR_CUDA, what, cudaGetErrorString(e__)); \
    } while (0)

#define VTF_CUDA(call, what)                                                                 \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) return vtf::fail(VTF_ERR_CUDA, what, cudaGetErrorString(e__)); \
    } while (0)

#define VTF_REQUIRE(cond, what)                                                              \
    do {                                                                                     \
        if (!(cond)) return vtf::fail(VTF_ERR_ARG, what, #cond);                             \
    } while (0)

int num_sms();

// Several labellings of the same rows handled in one pass (the k-range of cluster_faces, grouping.py:99-108):
// labelling L owns the centre / cluster slots [off[L], off[L+1]) of a concatenated [Ktot, d] array and the label
// vector labels[L * label_stride + row].  Bit L of `active` clear = labelling frozen (converged): skipped.
constexpr int kMaxSeg = 32;
struct Segs {
    int n;
    int off[kMaxSeg + 1];
    unsigned int active;
    long long label_stride;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit load that does not pollute L1 (X is read once per pass)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// out[e] = sum over parts in a fixed order (deterministic); the first lenA outputs go to outA,
// the rest to outB.  `static` so every translation unit owns a copy (no -rdc needed).
static __global__ void reduce_parts_kernel(const double* __restrict__ part, int nparts, int len,
                                           double* __restrict__ outA, double* __restrict__ outB, int lenA) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    double s = 0;
    for (int p = 0; p < nparts; ++p) s += part[(size_t)p * len + e];
    if (e < lenA) outA[e] = s; else outB[e - lenA] = s;
}

}  // namespace vtf
