// Row/column pre-passes of the grouping path: one coalesced, vectorised read of X each.
//   normalize_rows  <- sklearn normalize() inside cosine_distances (dupes.py:60, grouping.py:51)
//   col_stats/center_rows <- KMeans.fit set-up (SK/cluster/_kmeans.py:1473-1490)
//   gather_rows     <- np.delete (dupes.py:68)
// All HBM-bound: algorithmic bytes are N*d*4 read (+ N*d*4 written where a copy is produced).
#include "common.cuh"
#include <string.h>

namespace vtf {

thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};

int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

// ---------------------------------------------------------------------------------------
// warp per row: |x|^2 in fp32 (lanes own interleaved float4 columns), optional normalise.  The row stays
// in registers between the norm and the divide (DV float4 per lane, d = 128*DV), so X is read exactly once.
template <bool NORMALIZE, int DV>
__global__ void __launch_bounds__(256) row_norm_kernel(const float* __restrict__ X, int64_t N, int d,
                                                       float* __restrict__ Xn, float* __restrict__ sqn) {
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (; row < N; row += stride) {
        const float4* xr = reinterpret_cast<const float4*>(X + row * d);
        float4 xv[DV > 0 ? DV : 1];
        float s = 0.f;
        if (DV > 0) {
#pragma unroll
            for (int v = 0; v < DV; ++v) xv[v] = ldg_stream(xr + lane + 32 * v);
#pragma unroll
            for (int v = 0; v < DV; ++v) {
                s = fmaf(xv[v].x, xv[v].x, s); s = fmaf(xv[v].y, xv[v].y, s);
                s = fmaf(xv[v].z, xv[v].z, s); s = fmaf(xv[v].w, xv[v].w, s);
            }
        } else {
            for (int v = lane; v < nvec; v += 32) {
                float4 a = __ldg(xr + v);
                s = fmaf(a.x, a.x, s); s = fmaf(a.y, a.y, s); s = fmaf(a.z, a.z, s); s = fmaf(a.w, a.w, s);
            }
        }
        s = warp_sum(s);
        if (sqn && lane == 0) sqn[row] = s;
        if (NORMALIZE) {
            float nrm = sqrtf(s);
            // SK/preprocessing/_data.py:127-132: norms below 10*eps are replaced by 1
            if (nrm < 10.f * 1.1920929e-07f) nrm = 1.f;
            float4* outr = reinterpret_cast<float4*>(Xn + row * d);
            if (DV > 0) {
#pragma unroll
                for (int v = 0; v < DV; ++v) {
                    float4 a = xv[v];
                    a.x = __fdiv_rn(a.x, nrm); a.y = __fdiv_rn(a.y, nrm);
                    a.z = __fdiv_rn(a.z, nrm); a.w = __fdiv_rn(a.w, nrm);
                    outr[lane + 32 * v] = a;
                }
            } else {
                for (int v = lane; v < nvec; v += 32) {
                    float4 a = __ldg(xr + v);   // second touch comes from L1/L2
                    a.x = __fdiv_rn(a.x, nrm); a.y = __fdiv_rn(a.y, nrm);
                    a.z = __fdiv_rn(a.z, nrm); a.w = __fdiv_rn(a.w, nrm);
                    outr[v] = a;
                }
            }
        }
    }
}

// Encoder hand-off: FaceNet ends with F.normalize (encoders/facenet.py:153) and KMeans.fit starts with the
// column statistics of what comes out (SK/cluster/_kmeans.py:292, 1487).  For embeddings that are still un-normalised on the
// device both happen in ONE pass: warp per row (the row stays in registers between the norm and the divide), and every lane
// keeps fp64 column sums / sums of squares of the normalised values for the 4*DV columns it owns; warps and CTAs are
// combined in a fixed order (deterministic).  Saves the separate col_stats read of X.
template <int DV>
__global__ void __launch_bounds__(256) normalize_stats_kernel(const float* __restrict__ X, int64_t N, int d, float* __restrict__ Xn,
                                                             double* __restrict__ part) {
    extern __shared__ __align__(16) double sred[];     // [8 warps][2 * d]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double cs[DV][4], cq[DV][4];
#pragma unroll
    for (int v = 0; v < DV; ++v)
#pragma unroll
        for (int j = 0; j < 4; ++j) { cs[v][j] = 0.0; cq[v][j] = 0.0; }
    int64_t row = (int64_t)blockIdx.x * 8 + warp;
    const int64_t stride = (int64_t)gridDim.x * 8;
    for (; row < N; row += stride) {
        const float4* xr = reinterpret_cast<const float4*>(X + row * d);
        float4 xv[DV];
        float s = 0.f;
#pragma unroll
        for (int v = 0; v < DV; ++v) xv[v] = ldg_stream(xr + lane + 32 * v);
#pragma unroll
        for (int v = 0; v < DV; ++v) {
            s = fmaf(xv[v].x, xv[v].x, s); s = fmaf(xv[v].y, xv[v].y, s);
            s = fmaf(xv[v].z, xv[v].z, s); s = fmaf(xv[v].w, xv[v].w, s);
        }
        s = warp_sum(s);
        float nrm = sqrtf(s);
        if (nrm < 10.f * 1.1920929e-07f) nrm = 1.f;               // same guard as vtf_normalize_rows
        float4* outr = reinterpret_cast<float4*>(Xn + row * d);
#pragma unroll
        for (int v = 0; v < DV; ++v) {
            float4 a = xv[v];
            a.x = __fdiv_rn(a.x, nrm); a.y = __fdiv_rn(a.y, nrm); a.z = __fdiv_rn(a.z, nrm); a.w = __fdiv_rn(a.w, nrm);
            outr[lane + 32 * v] = a;
            cs[v][0] += a.x; cs[v][1] += a.y; cs[v][2] += a.z; cs[v][3] += a.w;
            cq[v][0] += (double)a.x * a.x; cq[v][1] += (double)a.y * a.y; cq[v][2] += (double)a.z * a.z; cq[v][3] += (double)a.w * a.w;
        }
    }
    double* mine = sred + (size_t)warp * 2 * d;
#pragma unroll
    for (int v = 0; v < DV; ++v)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = 4 * (lane + 32 * v) + j;
            mine[col] = cs[v][j];
            mine[d + col] = cq[v][j];
        }
    __syncthreads();
    double* out = part + (size_t)blockIdx.x * 2 * d;
    for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sred[(size_t)w * 2 * d + i];
        out[i] = t;
    }
}

template <bool NORMALIZE>
static int launch_row_norm(const float* X, int64_t N, int d, float* Xn, float* sqn, void* stream, const char* what) {
    const int grid = (int)min((int64_t)num_sms() * 8, (N + 7) / 8);
    const int dv = (d % 128 == 0 && d / 128 <= 8) ? d / 128 : 0;
#define VTF_RN_CASE(DVv) case DVv: VTF_LAUNCH((row_norm_kernel<NORMALIZE, DVv>), grid, 256, 0, stream, X, N, d, Xn, sqn); break;
    switch (dv) {
        VTF_RN_CASE(0) VTF_RN_CASE(1) VTF_RN_CASE(2) VTF_RN_CASE(3) VTF_RN_CASE(4)
        VTF_RN_CASE(5) VTF_RN_CASE(6) VTF_RN_CASE(7) VTF_RN_CASE(8)
    }
#undef VTF_RN_CASE
    VTF_CHECK_LAST(what);
    return VTF_OK;
}

// ---------------------------------------------------------------------------------------
// column sums / sums of squares in fp64: thread owns one float4 column, blocks stride rows.
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ X, int64_t N, int d,
                                                        double* __restrict__ part) {
    const int nvec = d >> 2;
    double* mypart = part + (size_t)blockIdx.x * 2 * d;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
        int64_t r = blockIdx.x;
        const int64_t step = gridDim.x;
        for (; r + 3 * step < N; r += 4 * step) {
            float4 a0 = ldg_stream(reinterpret_cast<const float4*>(X + r * d) + v);
            float4 a1 = ldg_stream(reinterpret_cast<const float4*>(X + (r + step) * d) + v);
            float4 a2 = ldg_stream(reinterpret_cast<const float4*>(X + (r + 2 * step) * d) + v);
            float4 a3 = ldg_stream(reinterpret_cast<const float4*>(X + (r + 3 * step) * d) + v);
#define VTF_ACC(a)                                                                     \
    s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;                                \
    q[0] += (double)a.x * a.x; q[1] += (double)a.y * a.y;                              \
    q[2] += (double)a.z * a.z; q[3] += (double)a.w * a.w;
            VTF_ACC(a0) VTF_ACC(a1) VTF_ACC(a2) VTF_ACC(a3)
        }
        for (; r < N; r += step) {
            float4 a0 = ldg_stream(reinterpret_cast<const float4*>(X + r * d) + v);
            VTF_ACC(a0)
        }
#undef VTF_ACC
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            mypart[4 * v + j] = s[j];
            mypart[d + 4 * v + j] = q[j];
        }
    }
}

__global__ void __launch_bounds__(256) center_kernel(const float4* __restrict__ X, int64_t nvec_total, int nvec,
                                                     const float4* __restrict__ mean, float4* __restrict__ Xc) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nvec_total; i += stride) {
        float4 a = ldg_stream(X + i);
        float4 m = __ldg(mean + (i % nvec));
        a.x -= m.x; a.y -= m.y; a.z -= m.z; a.w -= m.w;
        Xc[i] = a;
    }
}

// warp per row: Xc = X - mean, plus what the later stages ask of the centred matrix -- max |xc| (fp16 scales, fixed-point
// unit of the M-step sums) and max |xc_row|^2 (range of the norm pieces in the sweep's augmented operand) -- in the same
// pass; the row and the lane's slice of the mean stay in registers (DV float4 per lane, d = 128 * DV).
template <int DV>
__global__ void __launch_bounds__(256) center_stats_kernel(const float* __restrict__ X, int64_t N, int d, const float* __restrict__ mean,
                                                           float* __restrict__ Xc, float* __restrict__ sqn,
                                                           unsigned int* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    float4 mv[DV > 0 ? DV : 1];
    if (DV > 0) {
#pragma unroll
        for (int v = 0; v < DV; ++v) mv[v] = __ldg(reinterpret_cast<const float4*>(mean) + lane + 32 * v);
    }
    float amax = 0.f, smax = 0.f;
    int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (; row < N; row += stride) {
        const float4* xr = reinterpret_cast<const float4*>(X + row * d);
        float4* outr = reinterpret_cast<float4*>(Xc + row * d);
        float s = 0.f;
        if (DV > 0) {
            float4 xv[DV > 0 ? DV : 1];
#pragma unroll
            for (int v = 0; v < DV; ++v) xv[v] = ldg_stream(xr + lane + 32 * v);
#pragma unroll
            for (int v = 0; v < DV; ++v) {
                float4 a = xv[v];
                a.x -= mv[v].x; a.y -= mv[v].y; a.z -= mv[v].z; a.w -= mv[v].w;
                s = fmaf(a.x, a.x, s); s = fmaf(a.y, a.y, s); s = fmaf(a.z, a.z, s); s = fmaf(a.w, a.w, s);
                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
                outr[lane + 32 * v] = a;
            }
        } else {
            for (int v = lane; v < nvec; v += 32) {
                float4 a = ldg_stream(xr + v);
                const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + v);
                a.x -= m.x; a.y -= m.y; a.z -= m.z; a.w -= m.w;
                s = fmaf(a.x, a.x, s); s = fmaf(a.y, a.y, s); s = fmaf(a.z, a.z, s); s = fmaf(a.w, a.w, s);
                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
                outr[v] = a;
            }
        }
        s = warp_sum(s);
        smax = fmaxf(smax, s);
        if (sqn && lane == 0) sqn[row] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) {                                   // non-negative floats order like their bit patterns
        atomicMax(stats + 0, __float_as_uint(amax));
        atomicMax(stats + 1, __float_as_uint(smax));
    }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ X, int d, const int64_t* __restrict__ idx,
                                                          int64_t n_out, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (; row < n_out; row += stride) {
        const float4* src = reinterpret_cast<const float4*>(X + idx[row] * d);
        float4* dst = reinterpret_cast<float4*>(out + row * d);
        for (int v = lane; v < nvec; v += 32) dst[v] = ldg_stream(src + v);
    }
}

static int col_stats_grid() { return num_sms() * 8; }

}  // namespace vtf

using namespace vtf;

extern "C" {

int vtf_version(void) { return 100; }
const char* vtf_last_error(void) { return g_err; }
int64_t vtf_launch_count(void) { return (int64_t)g_launches.load(); }

int vtf_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(VTF_ERR_DEVICE, "vtf_device_check", cudaGetErrorString(e));
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        char buf[64];
        snprintf(buf, sizeof(buf), "compute capability %d.%d, need 10.x (B200)", major, minor);
        return fail(VTF_ERR_DEVICE, "vtf_device_check", buf);
    }
    return VTF_OK;
}

int vtf_normalize_rows(const float* X, int64_t N, int d, float* Xn, void* stream) {
    VTF_REQUIRE(X && Xn && N >= 0 && d > 0 && d % 4 == 0, "vtf_normalize_rows");
    if (N == 0) return VTF_OK;
    return launch_row_norm<true>(X, N, d, Xn, (float*)nullptr, stream, "vtf_normalize_rows");
}

int vtf_row_sqnorms(const float* X, int64_t N, int d, float* sqnorm, void* stream) {
    VTF_REQUIRE(X && sqnorm && N >= 0 && d > 0 && d % 4 == 0, "vtf_row_sqnorms");
    if (N == 0) return VTF_OK;
    return launch_row_norm<false>(X, N, d, (float*)nullptr, sqnorm, stream, "vtf_row_sqnorms");
}

size_t vtf_col_stats_ws_bytes(int d) { return (size_t)col_stats_grid() * 2 * (size_t)d * sizeof(double); }

int vtf_col_stats(const float* X, int64_t N, int d, double* col_sum, double* col_sumsq, void* ws, void* stream) {
    VTF_REQUIRE(X && col_sum && col_sumsq && ws && N >= 0 && d > 0 && d % 4 == 0, "vtf_col_stats");
    const int grid = col_stats_grid();
    int threads = min(256, ((d / 4 + 31) / 32) * 32);
    VTF_LAUNCH(col_stats_kernel, grid, threads, 0, stream, X, N, d, (double*)ws);
    VTF_CHECK_LAST("vtf_col_stats");
    VTF_LAUNCH(reduce_parts_kernel, (2 * d + 255) / 256, 256, 0, stream, (const double*)ws, grid, 2 * d, col_sum,
               col_sumsq, d);
    VTF_CHECK_LAST("vtf_col_stats/reduce");
    return VTF_OK;
}

int vtf_center_rows(const float* X, int64_t N, int d, const float* mean, float* Xc, void* stream) {
    VTF_REQUIRE(X && Xc && mean && N >= 0 && d > 0 && d % 4 == 0, "vtf_center_rows");
    if (N == 0) return VTF_OK;
    int64_t nvt = N * (d / 4);
    int grid = (int)min((int64_t)num_sms() * 16, (nvt + 255) / 256);
    VTF_LAUNCH(center_kernel, grid, 256, 0, stream, (const float4*)X, nvt, d / 4, (const float4*)mean, (float4*)Xc);
    VTF_CHECK_LAST("vtf_center_rows");
    return VTF_OK;
}

size_t vtf_normalize_col_stats_ws_bytes(int d) { return (size_t)num_sms() * 2 * (size_t)d * sizeof(double); }

int vtf_normalize_col_stats(const float* X, int64_t N, int d, float* Xn, double* col_sum, double* col_sumsq, void* ws, void* stream) {
    VTF_REQUIRE(X && Xn && col_sum && col_sumsq && ws && N >= 0 && d > 0 && d % 128 == 0 && d <= 1024, "vtf_normalize_col_stats (d % 128 == 0, d <= 1024)");
    const int grid = num_sms();
    const size_t smem = (size_t)8 * 2 * d * sizeof(double);
#define VTF_NS_CASE(DVv)                                                                                                       \
    case DVv:                                                                                                                  \
        VTF_CUDA(cudaFuncSetAttribute(normalize_stats_kernel<DVv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),    \
                 "vtf_normalize_col_stats/smem");                                                                              \
        VTF_LAUNCH((normalize_stats_kernel<DVv>), grid, 256, smem, stream, X, N, d, Xn, (double*)ws);                          \
        break;
    switch (d / 128) {
        VTF_NS_CASE(1) VTF_NS_CASE(2) VTF_NS_CASE(3) VTF_NS_CASE(4) VTF_NS_CASE(5) VTF_NS_CASE(6) VTF_NS_CASE(7) VTF_NS_CASE(8)
    }
#undef VTF_NS_CASE
    VTF_CHECK_LAST("vtf_normalize_col_stats");
    VTF_LAUNCH(reduce_parts_kernel, (2 * d + 255) / 256, 256, 0, stream, (const double*)ws, grid, 2 * d, col_sum, col_sumsq, d);
    VTF_CHECK_LAST("vtf_normalize_col_stats/reduce");
    return VTF_OK;
}

int vtf_center_rows_stats(const float* X, int64_t N, int d, const float* mean, float* Xc, float* sqn, float* stats,
                          void* stream) {
    VTF_REQUIRE(X && Xc && mean && stats && N >= 0 && d > 0 && d % 4 == 0, "vtf_center_rows_stats");
    VTF_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(float), (cudaStream_t)stream), "vtf_center_rows_stats");
    if (N == 0) return VTF_OK;
    const int grid = (int)min((int64_t)num_sms() * 8, (N + 7) / 8);
    const int dv = (d % 128 == 0 && d / 128 <= 8) ? d / 128 : 0;
#define VTF_CS_CASE(DVv) case DVv: VTF_LAUNCH((center_stats_kernel<DVv>), grid, 256, 0, stream, X, N, d, mean, Xc, sqn, (unsigned int*)stats); break;
    switch (dv) {
        VTF_CS_CASE(0) VTF_CS_CASE(1) VTF_CS_CASE(2) VTF_CS_CASE(3) VTF_CS_CASE(4)
        VTF_CS_CASE(5) VTF_CS_CASE(6) VTF_CS_CASE(7) VTF_CS_CASE(8)
    }
#undef VTF_CS_CASE
    VTF_CHECK_LAST("vtf_center_rows_stats");
    return VTF_OK;
}

int vtf_gather_rows(const float* X, int d, const int64_t* idx, int64_t n_out, float* X_out, void* stream) {
    VTF_REQUIRE(X && idx && X_out && n_out >= 0 && d > 0 && d % 4 == 0, "vtf_gather_rows");
    if (n_out == 0) return VTF_OK;
    int grid = (int)min((int64_t)num_sms() * 8, (n_out + 7) / 8);
    VTF_LAUNCH(gather_rows_kernel, grid, 256, 0, stream, X, d, idx, n_out, X_out);
    VTF_CHECK_LAST("vtf_gather_rows");
    return VTF_OK;
}

}  // extern "C"

// ---- device memory shared between ranks (CUDA IPC) ------------------------------------------------------------
extern "C" int vtf_shared_alloc(size_t bytes, void** ptr, unsigned char handle[64]) {
    VTF_REQUIRE(bytes > 0 && ptr && handle, "vtf_shared_alloc");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    void* p = nullptr;
    VTF_CUDA(cudaMalloc(&p, bytes), "vtf_shared_alloc/cudaMalloc");
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(VTF_ERR_CUDA, "vtf_shared_alloc", cudaGetErrorString(e)); }
    memcpy(handle, &h, 64);
    *ptr = p;
    return VTF_OK;
}

extern "C" int vtf_shared_open(const unsigned char handle[64], void** ptr) {
    VTF_REQUIRE(handle && ptr, "vtf_shared_open");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    VTF_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), "vtf_shared_open");
    return VTF_OK;
}

extern "C" int vtf_shared_zero(void* ptr, size_t bytes, void* stream) {
    VTF_REQUIRE(ptr && bytes > 0, "vtf_shared_zero");
    VTF_CUDA(cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream), "vtf_shared_zero");
    return VTF_OK;
}

extern "C" int vtf_shared_close(void* ptr) {
    VTF_REQUIRE(ptr, "vtf_shared_close");
    VTF_CUDA(cudaIpcCloseMemHandle(ptr), "vtf_shared_close");
    return VTF_OK;
}

extern "C" int vtf_shared_free(void* ptr) {
    VTF_REQUIRE(ptr, "vtf_shared_free");
    VTF_CUDA(cudaFree(ptr), "vtf_shared_free");
    return VTF_OK;
}
