// Rows-against-a-small-matrix kernels: the Lloyd E-step and nearest-reference classification.
//   lloyd_assign    <- SK/cluster/_k_means_lloyd.pyx:168-218 (SGEMM chunk + first-index argmin)
//   cosine_classify <- grouping.py:51-56 over SK/metrics/pairwise.py:1697-1752
// HBM-bound for the k / R these calls see (k <= 32, R ~ 32): one coalesced 128-bit read of each
// row, the small matrix staged once per CTA in shared memory, fp32 FMA dot products per lane
// and a shuffle reduction.  Algorithmic bytes: N*d*4 (+ N*R*4 for the distance log).
#include "common.cuh"

namespace vtf {

constexpr int kRowsPerWarp = 2;
constexpr int kMaxDV = 8;   // register-cached rows up to d = 1024

struct LloydEpi {
    int32_t* labels;                 // [n_seg][label_stride]
    unsigned long long* counts;      // [Ktot]
    unsigned long long* n_changed;   // [n_seg]
    Segs sg;
};
struct ClassifyEpi {
    float* dist;
    int64_t* label;
    float thr;       // < 0 or NaN: disabled
};

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    return acc;
}

// MODE 0: Lloyd (score = |c|^2 - 2 x.c), MODE 1: classify (score = clip(1 - x.c, 0, 2))
template <int MODE, int DV, bool C_SMEM>
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ X, int64_t N, int d,
                                                     const float* __restrict__ C, int k,
                                                     LloydEpi le, ClassifyEpi ce) {
    extern __shared__ __align__(16) float smem[];
    float* sC = smem;                          // [k*d] when C_SMEM
    float* sC2 = smem + (C_SMEM ? (size_t)k * d : 0);   // [k] squared norms (Lloyd)
    __shared__ unsigned int s_changed[kMaxSeg];
    unsigned int* s_counts = reinterpret_cast<unsigned int*>(sC2 + k);   // [k]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nvec = d >> 2;

    if (C_SMEM) {
        const float4* Cg = reinterpret_cast<const float4*>(C);
        float4* Cs = reinterpret_cast<float4*>(sC);
        for (int i = threadIdx.x; i < k * nvec; i += blockDim.x) Cs[i] = __ldg(Cg + i);
    }
    if (threadIdx.x < kMaxSeg) s_changed[threadIdx.x] = 0;
    for (int c = threadIdx.x; c < k; c += blockDim.x) s_counts[c] = 0;
    __syncthreads();
    if (MODE == 0) {
        // |c|^2 per centre, one warp per centre (fp32, as SK/cluster/_k_means_lloyd.pyx:99)
        for (int c = warp; c < k; c += nwarps) {
            const float4* cr = reinterpret_cast<const float4*>(C_SMEM ? sC + (size_t)c * d : C + (size_t)c * d);
            float s = 0.f;
            for (int v = lane; v < nvec; v += 32) { float4 a = cr[v]; s = dot4(a, a, s); }
            s = warp_sum(s);
            if (lane == 0) sC2[c] = s;
        }
        __syncthreads();
    }

    const int64_t warps_total = (int64_t)gridDim.x * nwarps;
    const int64_t step = warps_total * kRowsPerWarp;
    int64_t row0 = ((int64_t)blockIdx.x * nwarps + warp) * kRowsPerWarp;

    // Software pipeline: the rows (and, for Lloyd, the previous labels) of iteration i+1 are requested before
    // the dot products of iteration i start, so every warp keeps 4 KB of HBM reads in flight while it computes.
    float4 n0[DV > 0 ? DV : 1], n1[DV > 0 ? DV : 1];
    int old0 = -1, old1 = -1;
    auto prefetch = [&](int64_t r0) {
        if (r0 >= N) return;
        const float4* a0 = reinterpret_cast<const float4*>(X + r0 * d);
        const float4* a1 = reinterpret_cast<const float4*>(X + (r0 + 1 < N ? r0 + 1 : r0) * d);
        if (DV > 0) {
#pragma unroll
            for (int v = 0; v < DV; ++v) { n0[v] = ldg_stream(a0 + lane + 32 * v); n1[v] = ldg_stream(a1 + lane + 32 * v); }
        }
        if (MODE == 0 && lane < le.sg.n) {      // lane L keeps the previous label of labelling L
            old0 = le.labels[lane * le.sg.label_stride + r0];
            old1 = (r0 + 1 < N) ? le.labels[lane * le.sg.label_stride + r0 + 1] : -1;
        }
    };
    prefetch(row0);
    for (; row0 < N; row0 += step) {
        const bool has1 = (row0 + 1 < N);
        const float4* x0p = reinterpret_cast<const float4*>(X + row0 * d);
        const float4* x1p = reinterpret_cast<const float4*>(X + (has1 ? row0 + 1 : row0) * d);
        float4 x0[DV > 0 ? DV : 1], x1[DV > 0 ? DV : 1];
        if (DV > 0) {
#pragma unroll
            for (int v = 0; v < DV; ++v) { x0[v] = n0[v]; x1[v] = n1[v]; }
        }
        const int prev0 = old0, prev1 = old1;
        prefetch(row0 + step);
        float best0 = INFINITY, best1 = INFINITY;
        int arg0 = 0, arg1 = 0;
        float keep0 = 0.f, keep1 = 0.f;   // classify: lane (c & 31) keeps distance c for the coalesced store
        int seg = 0;                      // Lloyd: current labelling
        for (int c = 0; c < k; ++c) {
            if (MODE == 0 && c == le.sg.off[seg] && !((le.sg.active >> seg) & 1u)) {
                c = le.sg.off[seg + 1] - 1;   // frozen labelling: skip its centres
                ++seg;
                continue;
            }
            const float4* cr = reinterpret_cast<const float4*>(C_SMEM ? sC + (size_t)c * d : C + (size_t)c * d);
            float p0 = 0.f, p1 = 0.f, q0 = 0.f, q1 = 0.f;   // two partial sums per row: shorter FMA chains
            if (DV > 0) {
#pragma unroll
                for (int v = 0; v < DV; ++v) {
                    float4 cv = C_SMEM ? cr[lane + 32 * v] : __ldg(cr + lane + 32 * v);
                    if (v & 1) { q0 = dot4(x0[v], cv, q0); q1 = dot4(x1[v], cv, q1); }
                    else { p0 = dot4(x0[v], cv, p0); p1 = dot4(x1[v], cv, p1); }
                }
            } else {
                for (int v = lane; v < nvec; v += 32) {
                    float4 cv = C_SMEM ? cr[v] : __ldg(cr + v);
                    p0 = dot4(__ldg(x0p + v), cv, p0); p1 = dot4(__ldg(x1p + v), cv, p1);
                }
            }
            p0 = warp_sum(p0 + q0); p1 = warp_sum(p1 + q1);
            float s0, s1;
            if (MODE == 0) {
                const float c2 = sC2[c];
                s0 = fmaf(-2.f, p0, c2); s1 = fmaf(-2.f, p1, c2);
            } else {
                s0 = fminf(fmaxf(1.f - p0, 0.f), 2.f); s1 = fminf(fmaxf(1.f - p1, 0.f), 2.f);
                if ((c & 31) == lane) { keep0 = s0; keep1 = s1; }
                if ((c & 31) == 31 || c == k - 1) {
                    const int cb = c & ~31;
                    if (cb + lane <= c) {
                        ce.dist[row0 * k + cb + lane] = keep0;
                        if (has1) ce.dist[(row0 + 1) * k + cb + lane] = keep1;
                    }
                }
            }
            if (s0 < best0) { best0 = s0; arg0 = c; }   // strict '<': first index wins
            if (s1 < best1) { best1 = s1; arg1 = c; }
            if (MODE == 0 && c == le.sg.off[seg + 1] - 1) {
                // last centre of labelling `seg`: publish its labels, count changes, restart the argmin
                const int p0 = __shfl_sync(0xffffffffu, prev0, seg), p1 = __shfl_sync(0xffffffffu, prev1, seg);
                if (lane == 0) {
                    const int base = le.sg.off[seg];
                    unsigned int ch = (p0 != arg0 - base) ? 1u : 0u;
                    le.labels[seg * le.sg.label_stride + row0] = arg0 - base;
                    atomicAdd(&s_counts[arg0], 1u);
                    if (has1) {
                        ch += (p1 != arg1 - base) ? 1u : 0u;
                        le.labels[seg * le.sg.label_stride + row0 + 1] = arg1 - base;
                        atomicAdd(&s_counts[arg1], 1u);
                    }
                    if (ch) atomicAdd(&s_changed[seg], ch);
                }
                best0 = best1 = INFINITY;
                ++seg;
            }
        }
        if (lane == 0) {
            if (MODE == 0) {
            } else {
                const bool on = ce.thr >= 0.f;   // false for NaN as well
                ce.label[row0] = (on && best0 >= ce.thr) ? (int64_t)k : (int64_t)arg0;
                if (has1) ce.label[row0 + 1] = (on && best1 >= ce.thr) ? (int64_t)k : (int64_t)arg1;
            }
        }
    }
    if (MODE == 0) {
        __syncthreads();
        for (int c = threadIdx.x; c < k; c += blockDim.x)
            if (s_counts[c]) atomicAdd(le.counts + c, (unsigned long long)s_counts[c]);
        if (threadIdx.x < le.sg.n && s_changed[threadIdx.x])
            atomicAdd(le.n_changed + threadIdx.x, (unsigned long long)s_changed[threadIdx.x]);
    }
}

template <int MODE>
static int launch_rowdot(const float* X, int64_t N, int d, const float* C, int k, LloydEpi le, ClassifyEpi ce,
                         void* stream, const char* what) {
    const size_t small = (size_t)k * (sizeof(float) + sizeof(unsigned int));
    const size_t c_bytes = (size_t)k * d * sizeof(float);
    const bool c_smem = (c_bytes + small) <= 200 * 1024;
    const size_t smem = (c_smem ? c_bytes : 0) + small;
    const int dv = (d % 128 == 0 && d / 128 <= kMaxDV) ? d / 128 : 0;
    int blocks_per_sm = c_smem ? (int)max((size_t)1, min((size_t)3, (size_t)(220 * 1024) / (smem + 1024))) : 3;
    int64_t want = (N + 8 * kRowsPerWarp - 1) / (8 * kRowsPerWarp);
    int grid = (int)max((int64_t)1, min((int64_t)num_sms() * blocks_per_sm, want));

#define VTF_RD_CASE(DVv)                                                                               \
    case DVv:                                                                                          \
        if (c_smem) {                                                                                  \
            VTF_CUDA(cudaFuncSetAttribute(rowdot_kernel<MODE, DVv, true>,                              \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), what); \
            VTF_LAUNCH((rowdot_kernel<MODE, DVv, true>), grid, 256, smem, stream, X, N, d, C, k, le, ce); \
        } else {                                                                                       \
            VTF_LAUNCH((rowdot_kernel<MODE, DVv, false>), grid, 256, smem, stream, X, N, d, C, k, le, ce); \
        }                                                                                              \
        break;
    switch (dv) {
        VTF_RD_CASE(0) VTF_RD_CASE(1) VTF_RD_CASE(2) VTF_RD_CASE(3) VTF_RD_CASE(4)
        VTF_RD_CASE(5) VTF_RD_CASE(6) VTF_RD_CASE(7) VTF_RD_CASE(8)
    }
#undef VTF_RD_CASE
    VTF_CHECK_LAST(what);
    return VTF_OK;
}

// register-tiled fast path (rowtile.cu): 0 = launched, 1 = shape not covered, < 0 = error
int rowtile_lloyd(const float* X, int64_t N, int d, const float* C_all, const Segs& sg, int32_t* labels,
                  unsigned long long* counts, unsigned long long* n_changed, void* stream, const char* what);
int rowtile_classify(const float* X, int64_t N, int d, const float* R, int nR, float thr, float* dist, int64_t* label,
                     void* stream, const char* what);

static bool use_rowtile() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VTF_ROWTILE"); v = (e && atoi(e) == 0) ? 0 : 1; }    // VTF_ROWTILE=0: rowdot only
    return v == 1;
}

}  // namespace vtf

using namespace vtf;

extern "C" {

int vtf_lloyd_assign(const float* Xc, int64_t N, int d, const float* C, int k, int32_t* labels, int64_t* counts,
                     int64_t* n_changed, void* stream) {
    VTF_REQUIRE(Xc && C && labels && counts && n_changed && N >= 0 && d > 0 && d % 4 == 0 && k > 0,
                "vtf_lloyd_assign");
    if (N == 0) return VTF_OK;
    LloydEpi le{labels, reinterpret_cast<unsigned long long*>(counts), reinterpret_cast<unsigned long long*>(n_changed), {}};
    le.sg.n = 1; le.sg.off[0] = 0; le.sg.off[1] = k; le.sg.active = 1u; le.sg.label_stride = 0;
    if (use_rowtile()) {
        const int rc = rowtile_lloyd(Xc, N, d, C, le.sg, labels, le.counts, le.n_changed, stream, "vtf_lloyd_assign");
        if (rc <= 0) return rc;
    }
    ClassifyEpi ce{nullptr, nullptr, -1.f};
    return launch_rowdot<0>(Xc, N, d, C, k, le, ce, stream, "vtf_lloyd_assign");
}

int vtf_lloyd_assign_multi(const float* Xc, int64_t N, int d, const float* C_all, int n_lab, const int32_t* k_off_h,
                           uint32_t active, int32_t* labels, int64_t label_stride, int64_t* counts, int64_t* n_changed,
                           void* stream) {
    VTF_REQUIRE(Xc && C_all && k_off_h && labels && counts && n_changed && N >= 0 && d > 0 && d % 4 == 0 && n_lab > 0 &&
                    n_lab <= kMaxSeg && label_stride >= N,
                "vtf_lloyd_assign_multi (at most 32 labellings per call)");
    if (N == 0 || active == 0) return VTF_OK;
    LloydEpi le{labels, reinterpret_cast<unsigned long long*>(counts), reinterpret_cast<unsigned long long*>(n_changed), {}};
    le.sg.n = n_lab; le.sg.active = active; le.sg.label_stride = label_stride;
    for (int i = 0; i <= n_lab; ++i) le.sg.off[i] = k_off_h[i];
    VTF_REQUIRE(k_off_h[0] == 0 && k_off_h[n_lab] > 0, "vtf_lloyd_assign_multi: k_off must start at 0");
    if (use_rowtile()) {
        // the register-tiled kernel takes up to 32 active centres per pass: hand it the active labellings in runs of
        // sum(k) <= 32 (one pass over X each); whatever it cannot take (k > 32, odd d) stays with the warp-per-row kernel
        uint32_t left = active, run = 0;
        int run_k = 0;
        for (int L = 0; L <= n_lab; ++L) {
            const bool act = L < n_lab && ((active >> L) & 1u);
            const int kL = act ? k_off_h[L + 1] - k_off_h[L] : 0;
            if (run && (L == n_lab || (act && run_k + kL > 32))) {
                Segs sub = le.sg;
                sub.active = run;
                const int rc = rowtile_lloyd(Xc, N, d, C_all, sub, labels, le.counts, le.n_changed, stream, "vtf_lloyd_assign_multi");
                if (rc < 0) return rc;
                if (rc == 0) left &= ~run;
                run = 0; run_k = 0;
            }
            if (act && kL <= 32) { run |= 1u << L; run_k += kL; }
        }
        if (left == 0) return VTF_OK;
        le.sg.active = left;
    }
    ClassifyEpi ce{nullptr, nullptr, -1.f};
    return launch_rowdot<0>(Xc, N, d, C_all, k_off_h[n_lab], le, ce, stream, "vtf_lloyd_assign_multi");
}

int vtf_cosine_classify(const float* Xn, int64_t N, int d, const float* Rn, int R, float thr, float* dist,
                        int64_t* label, void* stream) {
    VTF_REQUIRE(Xn && Rn && dist && label && N >= 0 && d > 0 && d % 4 == 0 && R > 0, "vtf_cosine_classify");
    if (N == 0) return VTF_OK;
    if (use_rowtile()) {
        const int rc = rowtile_classify(Xn, N, d, Rn, R, thr, dist, label, stream, "vtf_cosine_classify");
        if (rc <= 0) return rc;
    }
    LloydEpi le{nullptr, nullptr, nullptr, {}};
    ClassifyEpi ce{dist, label, thr};
    return launch_rowdot<1>(Xn, N, d, Rn, R, le, ce, stream, "vtf_cosine_classify");
}

}  // extern "C"
