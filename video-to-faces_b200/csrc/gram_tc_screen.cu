// Two-stage duplicate marking for sm_100a (dupes.py:60-65): a single-MMA fp16 SCREEN of the lower triangle of Xn.Xn^T that
// keeps, per row, the best candidates among the earlier rows, followed by an exact fp32 re-evaluation of those candidates.
//
// gram_tc_dedup.cu reaches fp32-GEMM accuracy on EVERY pair by splitting the operands (hi.hi + hi.lo + lo.hi: three MMAs per
// K step) although only each row's minimum matters.  Here the sweep issues ONE MMA per K step on the rounded operand
// hi = fp16(s * xn) -- a third of the tensor work, half of the operand traffic -- and the epilogue keeps, for every row and
// each half of the column range it sees, the FOUR largest Gram values among the columns j < i (= the four smallest cosine
// distances; first index on equal values).  dedup_refine4_kernel then evaluates those up-to-eight pairs with exact fp32 dot
// products (the arithmetic of vtf_dedup_refine) and takes min / first argmin.
//
// What the screen may get wrong is bounded: |xn_i.xn_j - hi_i.hi_j / s^2| <= 2 * 2^-11 + 2^-22 for unit rows (fp16 keeps 11
// significant bits) plus the tensor core's truncating accumulation (~3e-6), together < 1.0e-3 =: eps.  A row's duplicate
// DECISION (min <= thr, dupes.py:65) is therefore exact unless its exact minimum over the evaluated candidates is above thr
// while an unseen column could still be below: 1 - g_unseen/s^2 - eps <= thr, g_unseen = the larger of the two halves'
// fourth-best values.  Such rows are rare (two neighbours inside a band of width eps right above the threshold); their 256-row
// units are flagged and re-swept by the split-fp16 kernel of gram_tc_dedup.cu (vtf_cosine_dedup_tc_units).
//
// Roles and pipeline as in gram_tc_dedup.cu: CTA pairs (cta_group::2), unit u = rows [256u, 256u+256) against column tiles
// 0..u, largest units first, dealt cyclically over ranks; warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue.
// Stage = {B_hi, A_hi} k-blocks of 16 KB, 6 stages; accumulator double-buffered in TMEM.  Algorithmic flops N^2*d.
#include "tc_common.cuh"

namespace vtf {
namespace tc {

struct ScreenParams {
    uint2* cand;               // [N][8] {float bits of g, column}: the four best of column half 0 (descending), then of half 1
    int64_t N;
    int64_t n_units;
    int64_t unit_begin;
    int64_t unit_stride;
    int n_kb;
};

constexpr int kSStages = 6;
constexpr int kSBlock = BM * BK * 2;            // 16 KB
constexpr int kSStageBytes = 2 * kSBlock;       // B_hi, A_hi

__global__ void __launch_bounds__(kThreads, 1)
dedup_screen_kernel(const __grid_constant__ CUtensorMap map_hi, const ScreenParams prm) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const uint32_t stage0 = base;
    const uint32_t bars = stage0 + kSStages * kSStageBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kSStages + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * kSStages) + 8u * s; };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * kSStages) + 16u + 8u * s; };
    const uint32_t tmem_slot = bars + 8u * (2 * kSStages) + 32u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
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

    const int64_t n_mine = (prm.n_units - prm.unit_begin + prm.unit_stride - 1) / prm.unit_stride;
    const int64_t pair0 = blockIdx.x / 2, n_pairs = gridDim.x / 2;
    auto unit_of = [&](int64_t pos) { return prm.n_units - 1 - (prm.unit_begin + pos * prm.unit_stride); };

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
                        if (leader) mbar_expect_tx(full_bar(s), (uint32_t)kSStageBytes * 2);
                        const uint32_t dst = stage0 + (uint32_t)s * kSStageBytes;
                        tma_load_2d_pair(&map_hi, full_bar(s), dst, kb * BK, brow);
                        tma_load_2d_pair(&map_hi, full_bar(s), dst + kSBlock, kb * BK, arow);
                        if (++s == kSStages) { s = 0; ph ^= 1u; }
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
                        const uint32_t sb = stage0 + (uint32_t)s * kSStageBytes;
                        const uint64_t b_hi = make_desc(sb), a_hi = make_desc(sb + kSBlock);
#pragma unroll
                        for (int k4 = 0; k4 < BK / UMMA_K; ++k4)
                            tc_mma_f16_pair(d_tmem, a_hi + (uint64_t)(2 * k4), b_hi + (uint64_t)(2 * k4), kIdescPair,
                                            (uint32_t)((kb | k4) != 0));
                        tc_commit_pair(empty_bar(s));
                        if (++s == kSStages) { s = 0; ph ^= 1u; }
                    }
                    tc_commit_pair(tfull_bar(as));
                    if (++as == 2) { as = 0; acc_ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue: thread = (row, column half)
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        int as = 0;
        uint32_t acc_ph = 0;
        for (int64_t pos = pair0; pos < n_mine; pos += n_pairs) {
            const int64_t u = unit_of(pos);
            const int64_t i = u * 256 + cta_rank * 128 + q * 32 + lane;     // this thread's row
            float g0 = -INFINITY, g1 = -INFINITY, g2 = -INFINITY, g3 = -INFINITY;      // the four largest Gram values over columns j < i
            uint32_t c0 = 0xffffffffu, c1 = 0xffffffffu, c2 = 0xffffffffu, c3 = 0xffffffffu;
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
                    float m = -INFINITY;
                    if (!diag) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float gj = (col0 + j < i) ? __uint_as_float(v[j]) : -INFINITY;   // only EARLIER rows (dupes.py:62)
                            v[j] = __float_as_uint(gj);
                            m = fmaxf(m, gj);
                        }
                    }
                    if (m > g3) {              // rare after the first tiles: something in this chunk enters the top four
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float gj = __uint_as_float(v[j]);
                            if (gj > g3) {     // insert and bubble up; strict comparisons: an equal later column never passes an earlier one
                                g3 = gj; c3 = (uint32_t)(col0 + j);
                                if (g3 > g2) { const float tg = g2; const uint32_t tcol = c2; g2 = g3; c2 = c3; g3 = tg; c3 = tcol; }
                                if (g2 > g1) { const float tg = g1; const uint32_t tcol = c1; g1 = g2; c1 = c2; g2 = tg; c2 = tcol; }
                                if (g1 > g0) { const float tg = g0; const uint32_t tcol = c0; g0 = g1; c0 = c1; g1 = tg; c1 = tcol; }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote<false>(tempty_bar(as), 0u);
                if (++as == 2) { as = 0; acc_ph ^= 1u; }
            }
            if (i < prm.N) {
                uint2* out = prm.cand + (size_t)i * 8 + 4 * half;
                out[0] = make_uint2(__float_as_uint(g0), c0);
                out[1] = make_uint2(__float_as_uint(g1), c1);
                out[2] = make_uint2(__float_as_uint(g2), c2);
                out[3] = make_uint2(__float_as_uint(g3), c3);
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// warp per owned row: exact fp32 cosine distance to the (up to eight) screened candidates, min and first argmin into key[];
// rows whose decision at `thr` the screen cannot guarantee flag their unit.
__global__ void __launch_bounds__(256) dedup_refine4_kernel(const float* __restrict__ Xn, int64_t N, int d, const uint2* __restrict__ cand,
                                                            float inv_s2, float thr, float eps, unsigned long long* __restrict__ key,
                                                            int* __restrict__ unit_flag, unsigned int* __restrict__ n_flagged,
                                                            int64_t row_stride_units, int64_t unit_rank) {
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * 8;
    const int64_t n_units = (N + 255) / 256;
    for (; row < N; row += stride) {
        if (((n_units - 1 - row / 256) % row_stride_units) != unit_rank) continue;
        const float4* a = reinterpret_cast<const float4*>(Xn + row * d);
        float best = INFINITY;
        uint32_t best_col = 0u;
        float g_unseen = -INFINITY;
        for (int c = 0; c < 8; ++c) {
            const uint2 e = cand[(size_t)row * 8 + c];
            if ((c & 3) == 3) g_unseen = fmaxf(g_unseen, __uint_as_float(e.x));     // a half's fourth best bounds everything it did not keep
            if (e.y == 0xffffffffu) continue;
            const float4* b = reinterpret_cast<const float4*>(Xn + (size_t)e.y * d);
            float s = 0.f;
            for (int v = lane; v < nvec; v += 32) {
                const float4 x = __ldg(a + v), y = __ldg(b + v);
                s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
            }
            s = warp_sum(s);
            const float dist = fminf(fmaxf(1.f - s, 0.f), 2.f);
            if (dist < best || (dist == best && e.y < best_col)) { best = dist; best_col = e.y; }
        }
        if (lane == 0 && best < INFINITY) {
            key[row] = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)best_col;
            // thr is NaN when no threshold applies: both comparisons are then false
            const float unseen_lb = 1.f - g_unseen * inv_s2 - eps;
            if (best > thr && unseen_lb <= thr) {
                if (atomicExch(unit_flag + row / 256, 1) == 0) atomicAdd(n_flagged, 1u);
            }
        }
    }
}

}  // namespace tc
}  // namespace vtf

using namespace vtf;

extern "C" {

int vtf_dedup_screen(const void* Hi, int64_t N, int64_t P, int d, int rank, int world, void* cand, void* stream) {
    VTF_REQUIRE(Hi && cand && N >= 0 && P >= N && P % 256 == 0 && d >= 64 && d <= 4096 && d % 64 == 0 && world > 0 && rank >= 0 &&
                    rank < world,
                "vtf_dedup_screen (P % 256 == 0, d % 64 == 0)");
    if (N <= 1) return VTF_OK;
    int rc = vtf_device_check();
    if (rc != VTF_OK) return rc;
    CUtensorMap mh;
    if ((rc = tc::make_operand_map(&mh, Hi, P, d, tc::BM, "vtf_dedup_screen")) != VTF_OK) return rc;
    tc::ScreenParams prm;
    prm.cand = (uint2*)cand;
    prm.N = N;
    prm.n_units = (N + 255) / 256;
    prm.unit_begin = rank;
    prm.unit_stride = world;
    prm.n_kb = d / tc::BK;
    if (prm.unit_begin >= prm.n_units) return VTF_OK;
    const int64_t n_mine = (prm.n_units - prm.unit_begin + prm.unit_stride - 1) / prm.unit_stride;
    const int grid = 2 * (int)min((int64_t)(num_sms() / 2), n_mine);
    const size_t smem = 1024 + (size_t)tc::kSStages * tc::kSStageBytes + 256;
    VTF_CUDA(cudaFuncSetAttribute(tc::dedup_screen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "vtf_dedup_screen/smem");
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
    VTF_CUDA(cudaLaunchKernelEx(&cfg, tc::dedup_screen_kernel, mh, prm), "vtf_dedup_screen/launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    VTF_CHECK_LAST("vtf_dedup_screen");
    return VTF_OK;
}

int vtf_dedup_refine4(const float* Xn, int64_t N, int d, int rank, int world, const void* cand, float scale, float thr, float eps,
                      uint64_t* key, int32_t* unit_flag, uint32_t* n_flagged, void* stream) {
    VTF_REQUIRE(Xn && cand && key && unit_flag && n_flagged && N >= 0 && d > 0 && d % 4 == 0 && world > 0 && rank >= 0 && rank < world &&
                    scale > 0.f && eps >= 0.f,
                "vtf_dedup_refine4");
    if (N <= 1) return VTF_OK;
    const int grid = (int)min((int64_t)num_sms() * 8, (N + 7) / 8);
    VTF_LAUNCH(tc::dedup_refine4_kernel, grid, 256, 0, stream, Xn, N, d, (const uint2*)cand, 1.f / (scale * scale), thr, eps,
               (unsigned long long*)key, unit_flag, n_flagged, (int64_t)world, (int64_t)rank);
    VTF_CHECK_LAST("vtf_dedup_refine4");
    return VTF_OK;
}

}  // extern "C"
