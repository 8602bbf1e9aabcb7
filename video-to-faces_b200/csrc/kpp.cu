// k-means++ seeding kernels (SK/cluster/_kmeans.py:180-278).
//   kpp_dist: distances from every row to T candidate centres, computed in fp64 and stored as
//             fp32 exactly like sklearn's upcast path (SK/metrics/pairwise.py:567-638), fused
//             with min(closest, .) (:262) and the per-candidate potential (:263).
//             HBM-bound: N*d*4 bytes read, T*N*4 written.
//   kpp_pick: searchsorted over the SEQUENTIAL fp32 cumsum of the closest distances (:250).
//             The running sum is one dependent chain of N fp32 additions, so one thread walks
//             it while the rest of the CTA streams the next chunk into shared memory; a
//             parallel scan would round differently and move candidates by a few rows.
#include "common.cuh"

namespace vtf {

constexpr int kMaxTrials = 8;

__global__ void __launch_bounds__(256) kpp_dist_kernel(const float* __restrict__ X, int64_t N, int d,
                                                       const float* __restrict__ cand, int T,
                                                       const float* __restrict__ closest, float* __restrict__ out,
                                                       double* __restrict__ part) {
    extern __shared__ __align__(16) float sc[];   // [T][d] candidates
    __shared__ double s_cc[kMaxTrials];
    __shared__ double s_pot[8][kMaxTrials];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = d >> 2;
    for (int i = threadIdx.x; i < T * nvec; i += blockDim.x)
        reinterpret_cast<float4*>(sc)[i] = __ldg(reinterpret_cast<const float4*>(cand) + i);
    __syncthreads();
    if (warp < T) {
        double s = 0.0;
        for (int j = lane; j < d; j += 32) { double c = (double)sc[warp * d + j]; s = fma(c, c, s); }
        s = warp_sum(s);
        if (lane == 0) s_cc[warp] = s;
    }
    __syncthreads();

    double pot[kMaxTrials];
#pragma unroll
    for (int t = 0; t < kMaxTrials; ++t) pot[t] = 0.0;

    int64_t row = (int64_t)blockIdx.x * 8 + warp;
    const int64_t stride = (int64_t)gridDim.x * 8;
    for (; row < N; row += stride) {
        const float4* xr = reinterpret_cast<const float4*>(X + row * d);
        double xx = 0.0, dot[kMaxTrials];
#pragma unroll
        for (int t = 0; t < kMaxTrials; ++t) dot[t] = 0.0;
        for (int v = lane; v < nvec; v += 32) {
            const float4 a = ldg_stream(xr + v);
            const double ax = a.x, ay = a.y, az = a.z, aw = a.w;
            xx = fma(ax, ax, xx); xx = fma(ay, ay, xx); xx = fma(az, az, xx); xx = fma(aw, aw, xx);
#pragma unroll
            for (int t = 0; t < kMaxTrials; ++t) {
                if (t < T) {
                    const float4 c = reinterpret_cast<const float4*>(sc + (size_t)t * d)[v];
                    dot[t] = fma(ax, (double)c.x, dot[t]); dot[t] = fma(ay, (double)c.y, dot[t]);
                    dot[t] = fma(az, (double)c.z, dot[t]); dot[t] = fma(aw, (double)c.w, dot[t]);
                }
            }
        }
        xx = warp_sum(xx);
        const float prev = closest ? __ldg(closest + row) : INFINITY;
#pragma unroll
        for (int t = 0; t < kMaxTrials; ++t) {
            if (t < T) {
                const double dt = warp_sum(dot[t]);
                // ((-2 x.c) + |c|^2) + |x|^2 in fp64, then fp32, then clamp (pairwise.py:632-636, 413)
                float dist = (float)((-2.0 * dt + s_cc[t]) + xx);
                dist = fmaxf(dist, 0.f);
                dist = fminf(dist, prev);
                if (lane == 0) {
                    out[(size_t)t * N + row] = dist;
                    pot[t] += (double)dist;
                }
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < kMaxTrials; ++t) s_pot[warp][t] = pot[t];
    }
    __syncthreads();
    if (threadIdx.x < T) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += s_pot[w][threadIdx.x];
        part[(size_t)blockIdx.x * T + threadIdx.x] = s;
    }
}

constexpr int kPickChunk = 4096;

__global__ void __launch_bounds__(128) kpp_pick_kernel(const float* __restrict__ closest, int64_t N,
                                                       const float* __restrict__ thr, int T,
                                                       const float* __restrict__ offset_in, float* __restrict__ total_out,
                                                       int64_t* __restrict__ idx) {
    __shared__ __align__(16) float buf[2][kPickChunk];
    __shared__ float sthr[kMaxTrials];
    if (threadIdx.x < T) { sthr[threadIdx.x] = thr[threadIdx.x]; idx[threadIdx.x] = -1; }
    const int64_t nchunks = (N + kPickChunk - 1) / kPickChunk;
    auto load = [&](int64_t c) {
        float* b = buf[c & 1];
        const int64_t base = c * kPickChunk;
        for (int i = threadIdx.x; i < kPickChunk; i += blockDim.x) {
            const int64_t g = base + i;
            b[i] = g < N ? __ldg(closest + g) : 0.f;
        }
    };
    if (nchunks > 0) load(0);
    __syncthreads();
    float s = offset_in ? *offset_in : 0.f;
    int t = 0;
    // thresholds already met by the incoming running sum belong to an earlier shard
    if (threadIdx.x == 0) { while (t < T && s >= sthr[t]) ++t; }
    float th = (threadIdx.x == 0 && t < T) ? sthr[t] : INFINITY;
    for (int64_t c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) load(c + 1);
        if (threadIdx.x == 0) {
            // Distances are >= 0, so the running sum never decreases: test one prefix per 16 elements
            // and replay the block element by element only when a threshold falls inside it.  The
            // chunk is zero-padded past N (s + 0 == s), so whole blocks are always safe.
            const float4* b4 = reinterpret_cast<const float4*>(buf[c & 1]);
            const int64_t base = c * kPickChunk;
            // software pipeline: the 4 x LDS.128 of block i+1 are in flight while the 16-add chain of
            // block i (the critical path: 16 x 4 cycles) runs
            float4 nq[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) nq[u] = b4[u];
#pragma unroll 1
            for (int i = 0; i < kPickChunk / 4; i += 8) {      // 32 elements per block: one compare per 32 adds
                float4 q[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) q[u] = nq[u];
                if (i + 8 < kPickChunk / 4) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) nq[u] = b4[i + 8 + u];
                }
                float e = s;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    e = __fadd_rn(e, q[u].x); e = __fadd_rn(e, q[u].y); e = __fadd_rn(e, q[u].z); e = __fadd_rn(e, q[u].w);
                }
                if (e >= th) {
                    float r = s;
                    const float* bs = buf[c & 1] + 4 * i;   // rare: replay the block element by element
#pragma unroll 1
                    for (int j = 0; j < 32; ++j) {
                        r = __fadd_rn(r, bs[j]);
                        while (r >= th) {      // searchsorted side='left': first i with cumsum[i] >= v
                            const int64_t g = base + 4 * i + j;
                            idx[t] = g < N ? g : -1;
                            ++t;
                            th = t < T ? sthr[t] : INFINITY;
                        }
                    }
                }
                s = e;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = s;
}

static int kpp_grid() { return num_sms() * 4; }

}  // namespace vtf

using namespace vtf;

extern "C" {

size_t vtf_kpp_ws_bytes(int T) { return (size_t)kpp_grid() * (size_t)(T > 0 ? T : 1) * sizeof(double); }

int vtf_kpp_dist(const float* Xc, int64_t N, int d, const float* cand, int T, const float* closest, float* out,
                 double* pots, void* ws, void* stream) {
    VTF_REQUIRE(Xc && cand && out && pots && ws && N >= 0 && d > 0 && d % 4 == 0 && T > 0 && T <= kMaxTrials,
                "vtf_kpp_dist");
    const size_t smem = (size_t)T * d * sizeof(float);
    VTF_REQUIRE(smem <= 64 * 1024, "vtf_kpp_dist: T*d too large");
    const int grid = kpp_grid();
    VTF_CUDA(cudaFuncSetAttribute(kpp_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), "vtf_kpp_dist");
    VTF_LAUNCH(kpp_dist_kernel, grid, 256, smem, stream, Xc, N, d, cand, T, closest, out, (double*)ws);
    VTF_CHECK_LAST("vtf_kpp_dist");
    VTF_LAUNCH(reduce_parts_kernel, 1, 32, 0, stream, (const double*)ws, grid, T, pots, pots, T);
    VTF_CHECK_LAST("vtf_kpp_dist/reduce");
    return VTF_OK;
}

int vtf_kpp_pick(const float* closest, int64_t N, const float* thr, int T, const float* offset_in, float* total_out,
                 int64_t* idx, void* stream) {
    VTF_REQUIRE(closest && thr && idx && N >= 0 && T > 0 && T <= kMaxTrials, "vtf_kpp_pick");
    VTF_LAUNCH(kpp_pick_kernel, 1, 128, 0, stream, closest, N, thr, T, offset_in, total_out, idx);
    VTF_CHECK_LAST("vtf_kpp_pick");
    return VTF_OK;
}

}  // extern "C"
