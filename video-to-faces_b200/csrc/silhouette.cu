// Silhouette support kernels around the Gram sweeps (gram_simt.cu, gram_tc.cu).
//   sil_pack          fp32 rows -> permuted, padded, power-of-two scaled fp16 operand + norms
//   absmax            max |x| (picks the power-of-two scale for fp16)
//   sil_finish_rows   per-row a, b, s from [N,k] cluster sums   (_unsupervised.py:191-199, 315-323)
//   sil_finish_cells  same from the per-cell sums of the tensor-core sweep
// HBM-bound, O(N*d) or O(N*n_cells) bytes; the O(N^2 d) work lives in the sweeps.
#include "common.cuh"

namespace vtf {

// Augmented operand ("the norms ride in the MMA"): 16 extra fp16 K-columns per row so that one more
// K=16 tcgen05.mma adds -m_p - m_q (m = |h|^2 / 2) to the Gram accumulator:
//   A flavour  [-m_hi, -m_mid, -m_lo,  C,  C,  C, 0 x 10]      B flavour  [C, C, C, -m_hi, -m_mid, -m_lo, 0 x 10]
// with m' = m * 2^-shift split into three fp16 pieces (33 bits >= fp32's 24) and C = 2^shift.
// Stored in the canonical no-swizzle K-major UMMA order so a plain tiled TMA copy lands it ready to use:
// element (p, k) at byte (p/8)*256 + (k/8)*128 + (p%8)*16 + (k%8)*2.
__device__ __forceinline__ void store_aug(__half* augA, __half* augB, int64_t p, float v, float cpow) {
    const __half hi = __float2half_rn(v);
    const float r1 = v - __half2float(hi);
    const __half mid = __float2half_rn(r1);
    const __half lo = __float2half_rn(r1 - __half2float(mid));
    const __half c = __float2half_rn(cpow), z = __float2half_rn(0.f);
    const __half nh = __hneg(hi), nm = __hneg(mid), nl = __hneg(lo);
    const size_t off = (size_t)(p >> 3) * 128 + (size_t)(p & 7) * 8;     // in halfs
    __half a0[8] = {nh, nm, nl, c, c, c, z, z};
    __half b0[8] = {c, c, c, nh, nm, nl, z, z};
    *reinterpret_cast<uint4*>(augA + off) = *reinterpret_cast<const uint4*>(a0);
    *reinterpret_cast<uint4*>(augB + off) = *reinterpret_cast<const uint4*>(b0);
    *reinterpret_cast<uint4*>(augA + off + 64) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(augB + off + 64) = make_uint4(0u, 0u, 0u, 0u);
}

// warp per output row p: Xh[p,:] = fp16_rn(scale * X[perm[p],:]); m_p = |Xh[p,:]|^2 / 2 computed from the
// ROUNDED values, so the distance is the exact distance between the rounded points.  Padding rows are zero
// vectors with an ALL-ZERO augmented block (no norm pieces and no constant either): every product that involves a
// padding row or column is exactly 0, so T = 0 and the distance is 0 without any clamp in the epilogue.
__global__ void __launch_bounds__(256) sil_pack_kernel(const float* __restrict__ X, int64_t N, int d,
                                                       const int64_t* __restrict__ perm, int64_t P, float scale,
                                                       float inv_pow, float cpow, float pad_m,
                                                       __half* __restrict__ Xh, __half* __restrict__ augA,
                                                       __half* __restrict__ augB, float* __restrict__ half_sqn) {
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (; p < P; p += stride) {
        const int64_t src = perm ? perm[p] : (p < N ? p : -1);
        uint2* out = reinterpret_cast<uint2*>(Xh + p * d);
        if (src < 0) {
            for (int v = lane; v < nvec; v += 32) out[v] = make_uint2(0u, 0u);
            if (lane == 0) {
                const size_t off = (size_t)(p >> 3) * 128 + (size_t)(p & 7) * 8;     // in halfs, as in store_aug
                *reinterpret_cast<uint4*>(augA + off) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(augB + off) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(augA + off + 64) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(augB + off + 64) = make_uint4(0u, 0u, 0u, 0u);
                if (half_sqn) half_sqn[p] = -pad_m;
            }
            continue;
        }
        const float4* xr = reinterpret_cast<const float4*>(X + src * d);
        float s = 0.f;
        for (int v = lane; v < nvec; v += 32) {
            const float4 a = ldg_stream(xr + v);
            const __half2 h01 = __floats2half2_rn(a.x * scale, a.y * scale);
            const __half2 h23 = __floats2half2_rn(a.z * scale, a.w * scale);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            s = fmaf(f01.x, f01.x, s); s = fmaf(f01.y, f01.y, s);
            s = fmaf(f23.x, f23.x, s); s = fmaf(f23.y, f23.y, s);
            uint2 o;
            o.x = *reinterpret_cast<const unsigned int*>(&h01);
            o.y = *reinterpret_cast<const unsigned int*>(&h23);
            out[v] = o;
        }
        s = warp_sum(s);
        if (lane == 0) {
            const float m = 0.5f * s;
            store_aug(augA, augB, p, m * inv_pow, cpow);
            if (half_sqn) half_sqn[p] = m;
        }
    }
}

// Augmented operand from given half squared norms m[p] (vtf_sil_diag: the tensor core's own |Xh_p|^2 / 2); rows with
// perm[p] < 0 (padding) get the all-zero block.
__global__ void __launch_bounds__(256) sil_aug_kernel(const float* __restrict__ m, const int64_t* __restrict__ perm, int64_t N, int64_t P,
                                                      float inv_pow, float cpow, __half* __restrict__ augA, __half* __restrict__ augB) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    if (perm ? perm[p] < 0 : p >= N) {
        const size_t off = (size_t)(p >> 3) * 128 + (size_t)(p & 7) * 8;
        *reinterpret_cast<uint4*>(augA + off) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(augB + off) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(augA + off + 64) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(augB + off + 64) = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    store_aug(augA, augB, p, m[p] * inv_pow, cpow);
}

__global__ void __launch_bounds__(256) absmax_kernel(const float4* __restrict__ X, int64_t nvec, unsigned int* __restrict__ out) {
    float m = 0.f;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nvec; i += stride) {
        const float4 a = ldg_stream(X + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));   // non-negative floats order as uints
}

__device__ __forceinline__ double sil_value(double a_sum, double b, long long n_own) {
    // a = intra / (n_own - 1); singleton clusters give 0 (nan_to_num, _unsupervised.py:323)
    if (n_own <= 1) return 0.0;
    const double a = a_sum / (double)(n_own - 1);
    const double m = fmax(a, b);
    if (!(m > 0.0) || !isfinite(b)) return 0.0;
    return (b - a) / m;
}

__device__ __forceinline__ void block_add(double v, double* out) {
    __shared__ double sred[8];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w];
        atomicAdd(out, s);
    }
}

__global__ void __launch_bounds__(256) sil_finish_rows_kernel(const double* __restrict__ rowsum, int k,
                                                              const int32_t* __restrict__ labels,
                                                              const int64_t* __restrict__ counts, int64_t row_begin,
                                                              int64_t row_end, double* __restrict__ out) {
    const int64_t i = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    if (i < row_end) {
        const int own = labels[i];
        double a_sum = 0.0, b = INFINITY;
        for (int c = 0; c < k; ++c) {
            const int64_t n = counts[c];
            if (n <= 0) continue;
            const double v = rowsum[i * k + c];
            if (c == own) a_sum = v; else b = fmin(b, v / (double)n);
        }
        s = sil_value(a_sum, b, counts[own]);
    }
    block_add(s, out);
}

// thread per sorted position p; shared S[c][t] collects the cells of each cluster
__global__ void __launch_bounds__(128) sil_finish_cells_kernel(const float* __restrict__ cellsum, int64_t P, int n_cells,
                                                               const int32_t* __restrict__ cell_to_cluster, int k,
                                                               const int64_t* __restrict__ perm,
                                                               const int32_t* __restrict__ labels,
                                                               const int64_t* __restrict__ counts, int64_t p_begin,
                                                               int64_t p_end, double* __restrict__ out) {
    extern __shared__ __align__(16) double S[];   // [k][128]
    const int t = threadIdx.x;
    const int64_t p = p_begin + (int64_t)blockIdx.x * blockDim.x + t;
    for (int c = 0; c < k; ++c) S[c * 128 + t] = 0.0;
    const bool live = p < p_end;
    const int64_t src = live ? perm[p] : -1;
    if (src >= 0) {
        for (int g = 0; g < n_cells; ++g) {
            const int c = __ldg(cell_to_cluster + g);
            S[c * 128 + t] += (double)__ldg(cellsum + (size_t)g * P + p);
        }
    }
    double s = 0.0;
    if (src >= 0) {
        const int own = labels[src];
        double a_sum = 0.0, b = INFINITY;
        for (int c = 0; c < k; ++c) {
            const int64_t n = counts[c];
            if (n <= 0) continue;
            const double v = S[c * 128 + t];
            if (c == own) a_sum = v; else b = fmin(b, v / (double)n);
        }
        s = sil_value(a_sum, b, counts[own]);
    }
    block_add(s, out);
}

}  // namespace vtf

using namespace vtf;

extern "C" {

int vtf_absmax(const float* X, int64_t N, int d, float* out, void* stream) {
    VTF_REQUIRE(X && out && N >= 0 && d > 0 && d % 4 == 0, "vtf_absmax");
    VTF_CUDA(cudaMemsetAsync(out, 0, sizeof(float), (cudaStream_t)stream), "vtf_absmax");
    if (N == 0) return VTF_OK;
    const int64_t nvec = N * (d / 4);
    const int grid = (int)min((int64_t)num_sms() * 8, (nvec + 255) / 256);
    VTF_LAUNCH(absmax_kernel, grid, 256, 0, stream, (const float4*)X, nvec, (unsigned int*)out);
    VTF_CHECK_LAST("vtf_absmax");
    return VTF_OK;
}

int vtf_sil_pack(const float* X, int64_t N, int d, const int64_t* perm, int64_t P, float scale, int shift, float pad_m,
                 void* Xh, void* aug, float* half_sqn, void* stream) {
    VTF_REQUIRE(X && Xh && aug && N >= 0 && P >= 0 && P % 8 == 0 && d > 0 && d % 4 == 0 && scale > 0.f && shift >= 0 &&
                    shift <= 15 && pad_m > 0.f,
                "vtf_sil_pack");
    if (P == 0) return VTF_OK;
    const int grid = (int)min((int64_t)num_sms() * 8, (P + 7) / 8);
    const float cpow = (float)(1 << shift);
    __half* augA = (__half*)aug;
    __half* augB = augA + (size_t)P * 16;
    VTF_LAUNCH(sil_pack_kernel, grid, 256, 0, stream, X, N, d, perm, P, scale, 1.f / cpow, cpow, pad_m, (__half*)Xh, augA,
               augB, half_sqn);
    VTF_CHECK_LAST("vtf_sil_pack");
    return VTF_OK;
}

int vtf_sil_aug(const float* m, const int64_t* perm, int64_t N, int64_t P, int shift, void* aug, void* stream) {
    VTF_REQUIRE(m && aug && N >= 0 && P >= 0 && P % 8 == 0 && shift >= 0 && shift <= 15, "vtf_sil_aug");
    if (P == 0) return VTF_OK;
    const float cpow = (float)(1 << shift);
    __half* augA = (__half*)aug;
    __half* augB = augA + (size_t)P * 16;
    VTF_LAUNCH(sil_aug_kernel, (unsigned)((P + 255) / 256), 256, 0, stream, m, perm, N, P, 1.f / cpow, cpow, augA, augB);
    VTF_CHECK_LAST("vtf_sil_aug");
    return VTF_OK;
}

int vtf_sil_finish_rows(const double* rowsum, int64_t N, int k, const int32_t* labels, const int64_t* counts,
                        int64_t row_begin, int64_t row_end, double* out_sum, void* stream) {
    VTF_REQUIRE(rowsum && labels && counts && out_sum && k > 0 && row_begin >= 0 && row_end <= N, "vtf_sil_finish_rows");
    if (row_end <= row_begin) return VTF_OK;
    const unsigned grid = (unsigned)((row_end - row_begin + 255) / 256);
    VTF_LAUNCH(sil_finish_rows_kernel, grid, 256, 0, stream, rowsum, k, labels, counts, row_begin, row_end, out_sum);
    VTF_CHECK_LAST("vtf_sil_finish_rows");
    return VTF_OK;
}

int vtf_sil_finish_cells(const float* cellsum, int64_t P, int n_cells, const int32_t* cell_to_cluster, int k,
                         const int64_t* perm, const int32_t* labels, const int64_t* counts, int64_t p_begin,
                         int64_t p_end, double* out_sum, void* stream) {
    VTF_REQUIRE(cellsum && cell_to_cluster && perm && labels && counts && out_sum && k > 0 && n_cells > 0 &&
                    p_begin >= 0 && p_end <= P,
                "vtf_sil_finish_cells");
    if (p_end <= p_begin) return VTF_OK;
    const size_t smem = (size_t)k * 128 * sizeof(double);
    VTF_REQUIRE(smem <= 200 * 1024, "vtf_sil_finish_cells: k too large");
    VTF_CUDA(cudaFuncSetAttribute(sil_finish_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
             "vtf_sil_finish_cells");
    const unsigned grid = (unsigned)((p_end - p_begin + 127) / 128);
    VTF_LAUNCH(sil_finish_cells_kernel, grid, 128, smem, stream, cellsum, P, n_cells, cell_to_cluster, k, perm, labels,
               counts, p_begin, p_end, out_sum);
    VTF_CHECK_LAST("vtf_sil_finish_cells");
    return VTF_OK;
}

}  // extern "C"
