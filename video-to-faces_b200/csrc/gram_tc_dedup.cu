// Tensor-core duplicate marking for sm_100a: the lower triangle of Xn.Xn^T on tcgen05 with a fused
// "distance -> running min / argmin over earlier rows" epilogue.
//
// Replaces dupes.py:60-65 at sizes where the fp32 FMA kernel of gram_simt.cu (36 TFLOP/s) is too slow:
// the reference materialises cosine_distances(X) (N*N fp32) plus an N*N fp64 np.tri mask and takes
// min/argmin per row.  Duplicate DECISIONS must match an fp32 SGEMM (hard part 2),
// so every operand is split into two fp16 pieces, s*x = hi + lo (22 significant bits), and the Gram
// value is accumulated in fp32 in TMEM from three fp16 MMAs per K step:
//      hi_i.hi_j  +  hi_i.lo_j  +  lo_i.hi_j          (lo.lo is below 2^-22 relative and dropped)
// Products of fp16 numbers are exact in fp32; what is left is the tensor core's truncating fp32 accumulation
// (~3e-6 relative, measured), so the sweep only FINDS each row's nearest earlier row and dedup_refine_kernel
// re-evaluates that one pair with an exact fp32 dot product before the threshold is applied.  3 x the MMA
// count of a plain fp16 Gram and still ~9 x faster than the fp32 pipe (200k x 768: 98 ms against 860 ms).
//
// CTA pairs (cta_group::2): unit u owns rows [256u, 256u+256) and sweeps column tiles 0..u (256 columns
// each); only tile u touches the diagonal and needs the j < i mask.  Units are dealt out largest first.
// Roles as in gram_tc.cu: warp 0 TMA producer (both CTAs), warp 1 MMA issuer (leader CTA), warps 2..9
// epilogue.  Stage = {B_hi, B_lo, A_hi, A_lo} k-blocks of 16 KB; 3 stages; accumulator double-buffered.
// Algorithmic flops: N^2*d (lower triangle); tensor-bound.
#include "tc_common.cuh"

namespace vtf {
namespace tc {

struct DedupParams {
    unsigned long long* key;   // [N] packed (bits(min dist) << 32) | argmin, pre-filled by vtf_dedup_init
    int64_t N;
    int64_t n_units;           // P / 256
    int64_t unit_begin;        // this rank takes units unit_begin, unit_begin + unit_stride, ... in DESCENDING row order
    int64_t unit_stride;
    int n_kb;                  // d / 64
    float inv_s2;              // 1 / scale^2
    const int32_t* unit_list;  // nullable: sweep exactly these units (the ones the single-MMA screen of gram_tc_screen.cu flagged)
    int64_t n_list;
};

constexpr int kDStages = 3;
constexpr int kDBlock = BM * BK * 2;            // 16 KB: 128 rows x 64 fp16
constexpr int kDStageBytes = 4 * kDBlock;       // B_hi, B_lo, A_hi, A_lo

__global__ void __launch_bounds__(kThreads, 1)
dedup_sweep_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const DedupParams prm) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const uint32_t stage0 = base;
    const uint32_t bars = stage0 + kDStages * kDStageBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kDStages + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * kDStages) + 8u * s; };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * kDStages) + 16u + 8u * s; };
    const uint32_t tmem_slot = bars + 8u * (2 * kDStages) + 32u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kDStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // this pair's units: positions pos = pair_id, pair_id + n_pairs, ... in this rank's list; unit = descending order
    const int64_t n_mine = prm.unit_list ? prm.n_list : (prm.n_units - prm.unit_begin + prm.unit_stride - 1) / prm.unit_stride;
    const int64_t pair0 = blockIdx.x / 2, n_pairs = gridDim.x / 2;
    auto unit_of = [&](int64_t pos) {
        return prm.unit_list ? (int64_t)prm.unit_list[pos] : prm.n_units - 1 - (prm.unit_begin + pos * prm.unit_stride);
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int64_t pos = pair0; pos < n_mine; pos += n_pairs) {
                const int64_t u = unit_of(pos);
                const int arow = (int)(u * 256 + cta_rank * 128);
                for (int64_t t = 0; t <= u; ++t) {
                    const int brow = (int)(t * 256 + cta_rank * 128);      // this CTA's half of the 256 B rows
                    for (int kb = 0; kb < prm.n_kb; ++kb) {
                        mbar_wait(empty_bar(s), ph ^ 1u);
                        if (leader) mbar_expect_tx(full_bar(s), (uint32_t)kDStageBytes * 2);
                        const uint32_t dst = stage0 + (uint32_t)s * kDStageBytes;
                        tma_load_2d_pair(&map_hi, full_bar(s), dst, kb * BK, brow);
                        tma_load_2d_pair(&map_lo, full_bar(s), dst + kDBlock, kb * BK, brow);
                        tma_load_2d_pair(&map_hi, full_bar(s), dst + 2 * kDBlock, kb * BK, arow);
                        tma_load_2d_pair(&map_lo, full_bar(s), dst + 3 * kDBlock, kb * BK, arow);
                        if (++s == kDStages) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (lane == 0 && leader) {
            int s = 0, as = 0;
            uint32_t ph = 0, acc_ph = 0;
            for (int64_t pos = pair0; pos < n_mine; pos += n_pairs) {
                const int64_t u = unit_of(pos);
                for (int64_t t = 0; t <= u; ++t) {
                    mbar_wait(tempty_bar(as), acc_ph ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)as * BN;
                    for (int kb = 0; kb < prm.n_kb; ++kb) {
                        mbar_wait(full_bar(s), ph);
                        tc_fence_after();
                        const uint32_t sb = stage0 + (uint32_t)s * kDStageBytes;
                        const uint64_t b_hi = make_desc(sb), b_lo = make_desc(sb + kDBlock);
                        const uint64_t a_hi = make_desc(sb + 2 * kDBlock), a_lo = make_desc(sb + 3 * kDBlock);
#pragma unroll
                        for (int k4 = 0; k4 < BK / UMMA_K; ++k4) {
                            const uint64_t o = (uint64_t)(2 * k4);
                            tc_mma_f16_pair(d_tmem, a_hi + o, b_hi + o, kIdescPair, (uint32_t)((kb | k4) != 0));
                            tc_mma_f16_pair(d_tmem, a_hi + o, b_lo + o, kIdescPair, 1u);
                            tc_mma_f16_pair(d_tmem, a_lo + o, b_hi + o, kIdescPair, 1u);
                        }
                        tc_commit_pair(empty_bar(s));
                        if (++s == kDStages) { s = 0; ph ^= 1u; }
                    }
                    tc_commit_pair(tfull_bar(as));
                    if (++as == 2) { as = 0; acc_ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (8 warps per CTA)
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        int as = 0;
        uint32_t acc_ph = 0;
        for (int64_t pos = pair0; pos < n_mine; pos += n_pairs) {
            const int64_t u = unit_of(pos);
            const int64_t i = u * 256 + cta_rank * 128 + q * 32 + lane;     // this thread's row
            float best = INFINITY;
            int64_t best_col = 0;
            for (int64_t t = 0; t <= u; ++t) {
                mbar_wait(tfull_bar(as), acc_ph);
                tc_fence_after();
                const bool diag = (t == u);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int64_t col0 = t * BN + half * 128 + c * 32;
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + (uint32_t)(as * BN + half * 128 + c * 32) + ((uint32_t)(q * 32) << 16);
                    VTF_TMEM_LD32(taddr, v);
                    VTF_TMEM_WAIT(v);
                    // D = clip(1 - S, 0, 2) with S = g / s^2 (SK/metrics/pairwise.py:1177-1180); min over the chunk
                    float m = INFINITY;
                    if (!diag) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float dj = fminf(fmaxf(fmaf(-__uint_as_float(v[j]), prm.inv_s2, 1.f), 0.f), 2.f);
                            v[j] = __float_as_uint(dj);
                            m = fminf(m, dj);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float dj = fminf(fmaxf(fmaf(-__uint_as_float(v[j]), prm.inv_s2, 1.f), 0.f), 2.f);
                            dj = (col0 + j < i) ? dj : INFINITY;             // only EARLIER rows (dupes.py:62)
                            v[j] = __float_as_uint(dj);
                            m = fminf(m, dj);
                        }
                    }
                    if (m < best) {            // strict: an equal distance further right never replaces the first
                        best = m;
                        int jj = 31;
#pragma unroll
                        for (int j = 31; j >= 0; --j) jj = (__uint_as_float(v[j]) == m) ? j : jj;
                        best_col = col0 + jj;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote<false>(tempty_bar(as), 0u);
                if (++as == 2) { as = 0; acc_ph ^= 1u; }
            }
            if (i < prm.N && best < INFINITY)
                atomicMin(prm.key + i, ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)best_col);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// warp per row: s*x = hi + lo with hi = fp16_rn(s*x), lo = fp16_rn(s*x - hi); rows >= N are zero
__global__ void __launch_bounds__(256) dedup_pack_kernel(const float* __restrict__ Xn, int64_t N, int d, int64_t P, float scale,
                                                         __half* __restrict__ Hi, __half* __restrict__ Lo) {
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * 8;
    for (; p < P; p += stride) {
        uint2* oh = reinterpret_cast<uint2*>(Hi + p * d);
        uint2* ol = reinterpret_cast<uint2*>(Lo + p * d);
        if (p >= N) {
            for (int v = lane; v < nvec; v += 32) { oh[v] = make_uint2(0u, 0u); ol[v] = make_uint2(0u, 0u); }
            continue;
        }
        const float4* xr = reinterpret_cast<const float4*>(Xn + p * d);
        for (int v = lane; v < nvec; v += 32) {
            const float4 a = ldg_stream(xr + v);
            const float sx = a.x * scale, sy = a.y * scale, sz = a.z * scale, sw = a.w * scale;
            const __half2 h01 = __floats2half2_rn(sx, sy), h23 = __floats2half2_rn(sz, sw);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(sx - f01.x, sy - f01.y), l23 = __floats2half2_rn(sz - f23.x, sw - f23.y);
            uint2 a_hi, a_lo;
            a_hi.x = *reinterpret_cast<const unsigned int*>(&h01); a_hi.y = *reinterpret_cast<const unsigned int*>(&h23);
            a_lo.x = *reinterpret_cast<const unsigned int*>(&l01); a_lo.y = *reinterpret_cast<const unsigned int*>(&l23);
            oh[v] = a_hi;
            ol[v] = a_lo;
        }
    }
}

// The tensor core adds each K=16 group into the fp32 accumulator with truncation (measured: S comes out low by
// ~0.5 ulp per MMA, about 3e-6 relative after 96 accumulations), which is fine for FINDING the nearest earlier
// row but not for the logged distance or a decision at the threshold.  So the winner of every row is
// re-evaluated exactly: one fp32 dot product x_i . x_argmin per row (warp per row, 2*N*d*4 bytes at most).
__global__ void __launch_bounds__(256) dedup_refine_kernel(const float* __restrict__ Xn, int64_t N, int d,
                                                           unsigned long long* __restrict__ key, int64_t row_stride_units,
                                                           int64_t unit_rank) {
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * 8;
    const int64_t n_units = (N + 255) / 256;
    for (; row < N; row += stride) {
        // only rows this rank owns carry a result (unit u = row/256 counted from the bottom, dealt out cyclically)
        if (((n_units - 1 - row / 256) % row_stride_units) != unit_rank) continue;
        const unsigned long long k = key[row];
        const int64_t j = (int64_t)(k & 0xffffffffull);
        if (__uint_as_float((unsigned)(k >> 32)) > 2.f) continue;       // row 0 / no earlier row: keeps 10000
        const float4* a = reinterpret_cast<const float4*>(Xn + row * d);
        const float4* b = reinterpret_cast<const float4*>(Xn + j * d);
        float s = 0.f;
        for (int v = lane; v < nvec; v += 32) {
            const float4 x = __ldg(a + v), y = __ldg(b + v);
            s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
        }
        s = warp_sum(s);
        if (lane == 0) {
            const float dist = fminf(fmaxf(1.f - s, 0.f), 2.f);
            key[row] = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)j;
        }
    }
}

}  // namespace tc
}  // namespace vtf

using namespace vtf;

extern "C" {

int vtf_dedup_pack(const float* Xn, int64_t N, int d, int64_t P, float scale, void* Hi, void* Lo, void* stream) {
    VTF_REQUIRE(Xn && Hi && Lo && N >= 0 && P >= N && P % 256 == 0 && d > 0 && d % 4 == 0 && scale > 0.f, "vtf_dedup_pack");
    if (P == 0) return VTF_OK;
    const int grid = (int)min((int64_t)num_sms() * 8, (P + 7) / 8);
    VTF_LAUNCH(tc::dedup_pack_kernel, grid, 256, 0, stream, Xn, N, d, P, scale, (__half*)Hi, (__half*)Lo);
    VTF_CHECK_LAST("vtf_dedup_pack");
    return VTF_OK;
}

static int dedup_tc_launch(const void* Hi, const void* Lo, int64_t N, int64_t P, int d, float scale, int rank, int world,
                           const int32_t* unit_list, int64_t n_list, uint64_t* key, void* stream);

int vtf_cosine_dedup_tc(const void* Hi, const void* Lo, int64_t N, int64_t P, int d, float scale, int rank, int world,
                        uint64_t* key, void* stream) {
    VTF_REQUIRE(Hi && Lo && key && N >= 0 && P >= N && P % 256 == 0 && d >= 64 && d <= 4096 && d % 64 == 0 && scale > 0.f &&
                    world > 0 && rank >= 0 && rank < world,
                "vtf_cosine_dedup_tc (P % 256 == 0, d % 64 == 0)");
    return dedup_tc_launch(Hi, Lo, N, P, d, scale, rank, world, nullptr, 0, key, stream);
}

int vtf_cosine_dedup_tc_units(const void* Hi, const void* Lo, int64_t N, int64_t P, int d, float scale, const int32_t* unit_list,
                              int64_t n_list, uint64_t* key, void* stream) {
    VTF_REQUIRE(Hi && Lo && key && unit_list && n_list >= 0 && N >= 0 && P >= N && P % 256 == 0 && d >= 64 && d <= 4096 &&
                    d % 64 == 0 && scale > 0.f,
                "vtf_cosine_dedup_tc_units (P % 256 == 0, d % 64 == 0)");
    if (n_list == 0) return VTF_OK;
    return dedup_tc_launch(Hi, Lo, N, P, d, scale, 0, 1, unit_list, n_list, key, stream);
}

static int dedup_tc_launch(const void* Hi, const void* Lo, int64_t N, int64_t P, int d, float scale, int rank, int world,
                           const int32_t* unit_list, int64_t n_list, uint64_t* key, void* stream) {
    if (N <= 1) return VTF_OK;
    int rc = vtf_device_check();
    if (rc != VTF_OK) return rc;
    CUtensorMap mh, ml;
    if ((rc = tc::make_operand_map(&mh, Hi, P, d, tc::BM, "vtf_cosine_dedup_tc")) != VTF_OK) return rc;
    if ((rc = tc::make_operand_map(&ml, Lo, P, d, tc::BM, "vtf_cosine_dedup_tc")) != VTF_OK) return rc;
    tc::DedupParams prm;
    prm.key = (unsigned long long*)key;
    prm.N = N;
    prm.n_units = (N + 255) / 256;          // units past the last real row hold padding only
    prm.unit_begin = rank;
    prm.unit_stride = world;
    prm.n_kb = d / tc::BK;
    prm.inv_s2 = 1.f / (scale * scale);
    prm.unit_list = unit_list;
    prm.n_list = n_list;
    if (!unit_list && prm.unit_begin >= prm.n_units) return VTF_OK;
    const int64_t n_mine = unit_list ? n_list : (prm.n_units - prm.unit_begin + prm.unit_stride - 1) / prm.unit_stride;
    const int grid = 2 * (int)min((int64_t)(num_sms() / 2), n_mine);
    const size_t smem = 1024 + (size_t)tc::kDStages * tc::kDStageBytes + 256;
    VTF_CUDA(cudaFuncSetAttribute(tc::dedup_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
             "vtf_cosine_dedup_tc/smem");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(tc::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VTF_CUDA(cudaLaunchKernelEx(&cfg, tc::dedup_sweep_kernel, mh, ml, prm), "vtf_cosine_dedup_tc/launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    VTF_CHECK_LAST("vtf_cosine_dedup_tc");
    return VTF_OK;
}

int vtf_dedup_refine(const float* Xn, int64_t N, int d, int rank, int world, uint64_t* key, void* stream) {
    VTF_REQUIRE(Xn && key && N >= 0 && d > 0 && d % 4 == 0 && world > 0 && rank >= 0 && rank < world, "vtf_dedup_refine");
    if (N <= 1) return VTF_OK;
    const int grid = (int)min((int64_t)num_sms() * 8, (N + 7) / 8);
    VTF_LAUNCH(tc::dedup_refine_kernel, grid, 256, 0, stream, Xn, N, d, (unsigned long long*)key, (int64_t)world, (int64_t)rank);
    VTF_CHECK_LAST("vtf_dedup_refine");
    return VTF_OK;
}

}  // extern "C"
