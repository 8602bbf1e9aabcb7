// Segmented (per-cluster) reductions over the rows of X.
//   cluster_sums       <- M-step accumulation, SK/cluster/_k_means_lloyd.pyx:215-218, and the
//                         per-cluster means of calinski_harabasz / davies_bouldin
//                         (SK/metrics/cluster/_unsupervised.py:394-399, 467-471)
//   cluster_dispersion <- KMeans.inertia_ (SK/cluster/_k_means_common.pyx:94-124), CH intra
//                         dispersion and DB intra distances
// HBM-bound (N*d*4 bytes per pass).  Accumulation is fp64 and DETERMINISTIC: every thread owns
// its output slots inside a CTA, CTAs write partials, and a fixed-order pass adds them -- the
// reference's OpenMP version reduces thread buffers in lock order (lloyd.pyx:143-152).
#include "common.cuh"

namespace vtf {

constexpr int kSliceDims = 256;   // dims per CTA slice == threads per CTA

// grid = (GX, ceil(d/256)); CTA (bx, by) adds rows [bx*rows_per, ...) restricted to dims
// [256*by, 256*by+256) into sacc[k][256] (shared, fp64, thread t owns column t of every cluster).
__global__ void __launch_bounds__(kSliceDims) cluster_sums_kernel(const float* __restrict__ X, int64_t N, int d,
                                                                  const int32_t* __restrict__ labels, int k, const Segs sg,
                                                                  int64_t rows_per, double* __restrict__ part) {
    extern __shared__ __align__(16) double sacc[];   // [k][256]
    const int t = threadIdx.x;
    const int dim = blockIdx.y * kSliceDims + t;
    const bool live = dim < d;
    for (int c = 0; c < k; ++c) sacc[c * kSliceDims + t] = 0.0;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per;
    const int64_t r1 = min(N, r0 + rows_per);
    const float* xp = X + (live ? dim : 0);
    int64_t r = r0;
    if (sg.n == 1) {
        for (; r + 8 <= r1; r += 8) {
            float v[8];
            int l[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                l[u] = __ldg(labels + r + u);
                v[u] = live ? __ldg(xp + (r + u) * d) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) sacc[l[u] * kSliceDims + t] += (double)v[u];
        }
        for (; r < r1; ++r) {
            const int l = __ldg(labels + r);
            const float v = live ? __ldg(xp + r * d) : 0.f;
            sacc[l * kSliceDims + t] += (double)v;
        }
    } else {
        // several labellings per pass: one read of the row feeds the slot of every active labelling
        for (; r + 4 <= r1; r += 4) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = live ? __ldg(xp + (r + u) * d) : 0.f;
            for (int L = 0; L < sg.n; ++L) {
                if (!((sg.active >> L) & 1u)) continue;
                const int32_t* lab = labels + L * sg.label_stride + r;
                const int base = sg.off[L];
#pragma unroll
                for (int u = 0; u < 4; ++u) sacc[(base + __ldg(lab + u)) * kSliceDims + t] += (double)v[u];
            }
        }
        for (; r < r1; ++r) {
            const float v = live ? __ldg(xp + r * d) : 0.f;
            for (int L = 0; L < sg.n; ++L) {
                if (!((sg.active >> L) & 1u)) continue;
                sacc[(sg.off[L] + __ldg(labels + L * sg.label_stride + r)) * kSliceDims + t] += (double)v;
            }
        }
    }
    if (live) {
        double* out = part + (size_t)blockIdx.x * k * d;
        for (int c = 0; c < k; ++c) out[(size_t)c * d + dim] = sacc[c * kSliceDims + t];
    }
}

// warp per row: d2 = |x - C64[label]|^2 in fp64; per-warp private accumulators in shared memory,
// fixed-order reduction over warps, then over CTAs (by reduce_parts_kernel).
__global__ void __launch_bounds__(256) cluster_dispersion_kernel(const float* __restrict__ X, int64_t N, int d,
                                                                 const int32_t* __restrict__ labels, int k, const Segs sg,
                                                                 const double* __restrict__ C64,
                                                                 double* __restrict__ part, float* __restrict__ far_d2) {
    extern __shared__ __align__(16) double sw[];   // [8 warps][2][k]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* mine = sw + (size_t)warp * 2 * k;
    for (int i = threadIdx.x; i < 8 * 2 * k; i += blockDim.x) sw[i] = 0.0;
    __syncthreads();
    const int nvec = d >> 2;
    int64_t row = (int64_t)blockIdx.x * 8 + warp;
    const int64_t stride = (int64_t)gridDim.x * 8;
    for (; row < N; row += stride) {
        const float4* xr = reinterpret_cast<const float4*>(X + row * d);
        for (int L = 0; L < sg.n; ++L) {
            if (!((sg.active >> L) & 1u)) continue;
            const int l = sg.off[L] + __ldg(labels + L * sg.label_stride + row);
            const double* cr = C64 + (size_t)l * d;
            double s = 0.0;
            for (int v = lane; v < nvec; v += 32) {
                const float4 a = L == 0 ? ldg_stream(xr + v) : __ldg(xr + v);   // later labellings re-read the row from L1/L2
                const double2 c0 = *reinterpret_cast<const double2*>(cr + 4 * v);
                const double2 c1 = *reinterpret_cast<const double2*>(cr + 4 * v + 2);
                double e;
                e = (double)a.x - c0.x; s = fma(e, e, s);
                e = (double)a.y - c0.y; s = fma(e, e, s);
                e = (double)a.z - c1.x; s = fma(e, e, s);
                e = (double)a.w - c1.y; s = fma(e, e, s);
            }
            s = warp_sum(s);
            if (lane == 0) {
                mine[l] += s;
                mine[k + l] += sqrt(s);
                if (far_d2 && L == 0) far_d2[row] = (float)s;
            }
        }
    }
    __syncthreads();
    double* out = part + (size_t)blockIdx.x * 2 * k;
    for (int i = threadIdx.x; i < 2 * k; i += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += sw[(size_t)w * 2 * k + i];
        out[i] = s;
    }
}

static int sums_gx(int d) {
    const int ds = (d + kSliceDims - 1) / kSliceDims;
    return max(1, num_sms() * 4 / ds);
}
static int disp_grid() { return num_sms() * 4; }

}  // namespace vtf

using namespace vtf;

extern "C" {

size_t vtf_cluster_sums_ws_bytes(int k, int d) {
    size_t a = (size_t)sums_gx(d) * (size_t)k * (size_t)d * sizeof(double);
    size_t b = (size_t)disp_grid() * 2 * (size_t)k * sizeof(double);
    return a > b ? a : b;
}

int vtf_cluster_sums(const float* X, int64_t N, int d, const int32_t* labels, int k, double* sums, void* ws,
                     void* stream) {
    VTF_REQUIRE(X && labels && sums && ws && N >= 0 && d > 0 && d % 4 == 0 && k > 0, "vtf_cluster_sums");
    const size_t smem = (size_t)k * kSliceDims * sizeof(double);
    VTF_REQUIRE(smem <= 200 * 1024, "vtf_cluster_sums: k too large (k <= 100)");
    const int gx = sums_gx(d);
    const int ds = (d + kSliceDims - 1) / kSliceDims;
    const int64_t rows_per = (N + gx - 1) / gx;
    VTF_CUDA(cudaFuncSetAttribute(cluster_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
             "vtf_cluster_sums");
    Segs sg = {};
    sg.n = 1; sg.off[0] = 0; sg.off[1] = k; sg.active = 1u; sg.label_stride = 0;
    VTF_LAUNCH(cluster_sums_kernel, dim3(gx, ds), kSliceDims, smem, stream, X, N, d, labels, k, sg, rows_per, (double*)ws);
    VTF_CHECK_LAST("vtf_cluster_sums");
    const int len = k * d;
    VTF_LAUNCH(reduce_parts_kernel, (len + 255) / 256, 256, 0, stream, (const double*)ws, gx, len, sums, sums, len);
    VTF_CHECK_LAST("vtf_cluster_sums/reduce");
    return VTF_OK;
}

static int fill_segs(Segs& sg, int n_lab, const int32_t* k_off_h, uint32_t active, int64_t label_stride, const char* what) {
    if (n_lab <= 0 || n_lab > kMaxSeg || !k_off_h || k_off_h[0] != 0) return fail(VTF_ERR_ARG, what, "1..32 labellings, k_off[0] == 0");
    sg.n = n_lab; sg.active = active; sg.label_stride = label_stride;
    for (int i = 0; i <= n_lab; ++i) sg.off[i] = k_off_h[i];
    return VTF_OK;
}

int vtf_cluster_sums_multi(const float* X, int64_t N, int d, const int32_t* labels, int64_t label_stride, int n_lab,
                           const int32_t* k_off_h, uint32_t active, double* sums, void* ws, void* stream) {
    VTF_REQUIRE(X && labels && sums && ws && N >= 0 && d > 0 && d % 4 == 0 && label_stride >= N, "vtf_cluster_sums_multi");
    Segs sg = {};
    int rc = fill_segs(sg, n_lab, k_off_h, active, label_stride, "vtf_cluster_sums_multi");
    if (rc != VTF_OK) return rc;
    const int k = k_off_h[n_lab];
    const size_t smem = (size_t)k * kSliceDims * sizeof(double);
    VTF_REQUIRE(smem <= 200 * 1024, "vtf_cluster_sums_multi: sum of k too large (<= 100 per call)");
    const int gx = sums_gx(d);
    const int ds = (d + kSliceDims - 1) / kSliceDims;
    const int64_t rows_per = (N + gx - 1) / gx;
    VTF_CUDA(cudaFuncSetAttribute(cluster_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
             "vtf_cluster_sums_multi");
    VTF_LAUNCH(cluster_sums_kernel, dim3(gx, ds), kSliceDims, smem, stream, X, N, d, labels, k, sg, rows_per, (double*)ws);
    VTF_CHECK_LAST("vtf_cluster_sums_multi");
    const int len = k * d;
    VTF_LAUNCH(reduce_parts_kernel, (len + 255) / 256, 256, 0, stream, (const double*)ws, gx, len, sums, sums, len);
    VTF_CHECK_LAST("vtf_cluster_sums_multi/reduce");
    return VTF_OK;
}

int vtf_cluster_dispersion_multi(const float* X, int64_t N, int d, const int32_t* labels, int64_t label_stride, int n_lab,
                                 const int32_t* k_off_h, uint32_t active, const double* C64, double* sq, double* rt, void* ws,
                                 void* stream) {
    VTF_REQUIRE(X && labels && C64 && sq && rt && ws && N >= 0 && d > 0 && d % 4 == 0 && label_stride >= N,
                "vtf_cluster_dispersion_multi");
    Segs sg = {};
    int rc = fill_segs(sg, n_lab, k_off_h, active, label_stride, "vtf_cluster_dispersion_multi");
    if (rc != VTF_OK) return rc;
    const int k = k_off_h[n_lab];
    const size_t smem = (size_t)8 * 2 * k * sizeof(double);
    VTF_REQUIRE(smem <= 48 * 1024, "vtf_cluster_dispersion_multi: sum of k too large");
    const int grid = disp_grid();
    VTF_LAUNCH(cluster_dispersion_kernel, grid, 256, smem, stream, X, N, d, labels, k, sg, C64, (double*)ws, (float*)nullptr);
    VTF_CHECK_LAST("vtf_cluster_dispersion_multi");
    VTF_LAUNCH(reduce_parts_kernel, (2 * k + 255) / 256, 256, 0, stream, (const double*)ws, grid, 2 * k, sq, rt, k);
    VTF_CHECK_LAST("vtf_cluster_dispersion_multi/reduce");
    return VTF_OK;
}

int vtf_cluster_dispersion(const float* X, int64_t N, int d, const int32_t* labels, int k, const double* C64,
                           double* sq, double* rt, float* far_d2, void* ws, void* stream) {
    VTF_REQUIRE(X && labels && C64 && sq && rt && ws && N >= 0 && d > 0 && d % 4 == 0 && k > 0,
                "vtf_cluster_dispersion");
    const size_t smem = (size_t)8 * 2 * k * sizeof(double);
    VTF_REQUIRE(smem <= 48 * 1024, "vtf_cluster_dispersion: k too large");
    const int grid = disp_grid();
    Segs sg = {};
    sg.n = 1; sg.off[0] = 0; sg.off[1] = k; sg.active = 1u; sg.label_stride = 0;
    VTF_LAUNCH(cluster_dispersion_kernel, grid, 256, smem, stream, X, N, d, labels, k, sg, C64, (double*)ws, far_d2);
    VTF_CHECK_LAST("vtf_cluster_dispersion");
    VTF_LAUNCH(reduce_parts_kernel, (2 * k + 255) / 256, 256, 0, stream, (const double*)ws, grid, 2 * k, sq, rt, k);
    VTF_CHECK_LAST("vtf_cluster_dispersion/reduce");
    return VTF_OK;
}

}  // extern "C"

namespace vtf {

// one CTA; k*d is tiny next to N*d
__global__ void __launch_bounds__(256) lloyd_update_kernel(const double* __restrict__ sums, const long long* __restrict__ counts,
                                                           const long long* __restrict__ n_changed, int k, int d,
                                                           const float* __restrict__ C_old, float* __restrict__ C_new,
                                                           double* __restrict__ C64, double* __restrict__ stats) {
    __shared__ double sred[8];
    int amax = 0;
    long long best = -1;
    for (int c = 0; c < k; ++c) { const long long n = counts[c]; if (n > best) { best = n; amax = c; } }   // np.argmax: first max
    double shift = 0.0;
    for (int e = threadIdx.x; e < k * d; e += blockDim.x) {
        const int c = e / d, j = e - c * d;
        const long long n = counts[c];
        const int src = n > 0 ? c : amax;
        const long long ns = n > 0 ? n : best;
        const float v = ns > 0 ? (float)(sums[(size_t)src * d + j] / (double)ns) : C_old[e];
        const double df = (double)v - (double)C_old[e];
        shift = fma(df, df, shift);
        C_new[e] = v;
        if (C64) C64[e] = (double)v;
    }
    shift = warp_sum(shift);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = shift;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += sred[w];
        stats[0] = s;
        stats[1] = n_changed ? (double)*n_changed : 0.0;
    }
    for (int c = threadIdx.x; c < k; c += blockDim.x) stats[2 + c] = (double)counts[c];
}

}  // namespace vtf

namespace vtf {
// all labellings in one launch: labelling L owns rows [off[L], off[L+1]) of sums / C and stats[L * stride .. ]
__global__ void __launch_bounds__(256) lloyd_update_multi_kernel(const double* __restrict__ sums, const long long* __restrict__ counts,
                                                                 const long long* __restrict__ n_changed, const Segs sg, int d,
                                                                 const float* __restrict__ C_old, float* __restrict__ C_new,
                                                                 double* __restrict__ C64, double* __restrict__ stats, int stride) {
    __shared__ double sred[8];
    for (int L = 0; L < sg.n; ++L) {
        if (!((sg.active >> L) & 1u)) continue;
        const int base = sg.off[L], k = sg.off[L + 1] - base;
        int amax = 0;
        long long best = -1;
        for (int c = 0; c < k; ++c) { const long long n = counts[base + c]; if (n > best) { best = n; amax = c; } }
        double shift = 0.0;
        for (int e = threadIdx.x; e < k * d; e += blockDim.x) {
            const int c = e / d, j = e - c * d;
            const long long n = counts[base + c];
            const int src = n > 0 ? c : amax;
            const long long ns = n > 0 ? n : best;
            const size_t ge = (size_t)(base + c) * d + j;
            const float v = ns > 0 ? (float)(sums[(size_t)(base + src) * d + j] / (double)ns) : C_old[ge];
            const double df = (double)v - (double)C_old[ge];
            shift = fma(df, df, shift);
            C_new[ge] = v;
            if (C64) C64[ge] = (double)v;
        }
        shift = warp_sum(shift);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = shift;
        __syncthreads();
        double* st = stats + (size_t)L * stride;
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += sred[w];
            st[0] = s;
            st[1] = (double)n_changed[L];
        }
        for (int c = threadIdx.x; c < k; c += blockDim.x) st[2 + c] = (double)counts[base + c];
    }
}
}  // namespace vtf

extern "C" int vtf_lloyd_update_multi(const double* sums, const int64_t* counts, const int64_t* n_changed, int n_lab,
                                      const int32_t* k_off_h, uint32_t active, int d, const float* C_old, float* C_new,
                                      double* C64, double* stats, int stats_stride, void* stream) {
    VTF_REQUIRE(sums && counts && n_changed && C_old && C_new && stats && d > 0 && stats_stride >= 2, "vtf_lloyd_update_multi");
    vtf::Segs sg = {};
    int rc = fill_segs(sg, n_lab, k_off_h, active, 0, "vtf_lloyd_update_multi");
    if (rc != VTF_OK) return rc;
    VTF_LAUNCH(vtf::lloyd_update_multi_kernel, 1, 256, 0, stream, sums, (const long long*)counts, (const long long*)n_changed, sg, d,
               C_old, C_new, C64, stats, stats_stride);
    VTF_CHECK_LAST("vtf_lloyd_update_multi");
    return VTF_OK;
}

extern "C" int vtf_lloyd_update(const double* sums, const int64_t* counts, const int64_t* n_changed, int k, int d,
                                const float* C_old, float* C_new, double* C64, double* stats, void* stream) {
    VTF_REQUIRE(sums && counts && C_old && C_new && stats && k > 0 && d > 0, "vtf_lloyd_update");
    VTF_LAUNCH(vtf::lloyd_update_kernel, 1, 256, 0, stream, sums, (const long long*)counts, (const long long*)n_changed, k, d,
               C_old, C_new, C64, stats);
    VTF_CHECK_LAST("vtf_lloyd_update");
    return VTF_OK;
}

namespace vtf {

// ---- contingency table of two labellings (sklearn.metrics.cluster.contingency_matrix, the input of rand_score's pair
// confusion matrix, SK/metrics/cluster/_supervised.py:195-271): out[a[i] * nb + b[i]] += 1.  Shared-memory histogram per
// CTA when the table fits (na*nb <= kContBins), global atomics otherwise; labels outside [0,na) x [0,nb) are counted in
// *bad so that the host can refuse them.
constexpr int kContBins = 8192;
__global__ void __launch_bounds__(256) contingency_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t N,
                                                          int na, int nb, unsigned long long* __restrict__ out,
                                                          unsigned long long* __restrict__ bad) {
    __shared__ unsigned int hist[kContBins];
    const int bins = na * nb;
    const bool local = bins <= kContBins;
    if (local) { for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = 0u; __syncthreads(); }
    unsigned int nbad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = __ldg(a + i), y = __ldg(b + i);
        if ((unsigned)x >= (unsigned)na || (unsigned)y >= (unsigned)nb) { ++nbad; continue; }
        if (local) atomicAdd(&hist[x * nb + y], 1u);
        else atomicAdd(&out[(size_t)x * nb + y], 1ull);
    }
    if (nbad) atomicAdd(bad, (unsigned long long)nbad);
    if (local) {
        __syncthreads();
        for (int i = threadIdx.x; i < bins; i += blockDim.x) { const unsigned int v = hist[i]; if (v) atomicAdd(&out[i], (unsigned long long)v); }
    }
}

}  // namespace vtf

extern "C" int vtf_contingency(const int32_t* a, const int32_t* b, int64_t N, int na, int nb, int64_t* table, int64_t* bad,
                               void* stream) {
    VTF_REQUIRE(a && b && table && bad && N >= 0 && na > 0 && nb > 0 && (int64_t)na * nb <= (1 << 26), "vtf_contingency");
    VTF_CUDA(cudaMemsetAsync(table, 0, (size_t)na * nb * sizeof(int64_t), (cudaStream_t)stream), "vtf_contingency");
    VTF_CUDA(cudaMemsetAsync(bad, 0, sizeof(int64_t), (cudaStream_t)stream), "vtf_contingency");
    if (N == 0) return VTF_OK;
    // per-CTA rows cap at 2^32-1 increments per bin: N / grid < 2^32 always holds for grid >= 1 and N < 2^40
    const int64_t want = (N + 255) / 256;
    const int grid = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
    VTF_LAUNCH(contingency_kernel, grid, 256, 0, stream, a, b, N, na, nb, (unsigned long long*)table, (unsigned long long*)bad);
    VTF_CHECK_LAST("vtf_contingency");
    return VTF_OK;
}
