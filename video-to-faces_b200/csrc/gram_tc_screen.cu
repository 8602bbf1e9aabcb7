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
// Roles as in gram_tc.cu: CTA pairs (cta_group::2), one 256 x 256 MMA tile per pair and step; warp 0 TMA producer, warp 1 MMA
// issuer, warps 2..9 epilogue; accumulator double-buffered in TMEM.  For d <= 512 each CTA keeps its 128 x d A block RESIDENT in
// shared memory (loaded once per work item) and streams only its half of every B k-block (16 KB, 6 ring stages); wider rows
// stream A with B (32 KB stages).
//
// Work.  Unit u = rows [256u, 256u+256) needs the column tiles 0..u.  Round 2's first version dealt whole units out round-robin,
// largest first: at 50k rows (196 units on 74 pairs) the pair that drew units 0, 74, 148 swept 366 tiles against an ideal of
// 261 (tensor pipe 38 % of elapsed), and at 4M rows the 74 pairs drifted apart along their units until every B tile came from
// DRAM again (6 GB of operand against 126 MB of L2: 914 TFLOP/s where the silhouette sweep reaches 1300+).  Now
//   * the triangle is cut into COLUMN CHUNKS of at most 48 MB of operand, one launch per chunk (stream order = the order the
//     candidate lists are continued in): all pairs of a launch read the same chunk, which comes from DRAM once; a row's lists
//     travel from chunk to chunk through `cand` (loaded at the start of an item, stored at its end);
//   * inside a launch the items (unit x chunk) are CLAIMED from a counter, longest first (greedy list scheduling): a pair that
//     finishes early takes the next item, so the launch ends within one short item of the ideal.
// Algorithmic flops N^2*d, executed once.
#include <stdlib.h>
#include "tc_common.cuh"

namespace vtf {
namespace tc {

struct ScreenParams {
    uint2* cand;               // [N][8] {float bits of g, column}: the four best of column half 0 (descending), then of half 1
    int64_t N;
    int64_t n_units;
    int64_t unit_begin;        // rank
    int64_t unit_stride;       // world
    int n_kb;
    int t_lo, t_hi;            // column tiles [t_lo, t_hi) of this launch
    int first;                 // 1: the lists start empty; 0: continue the lists in `cand`
    long long n_items;         // this rank's units u >= t_lo
    unsigned int* work_counter;   // zero at launch; claimed with atomicAdd
};

#ifndef VTF_SCREEN_RELAXED
#define VTF_SCREEN_RELAXED 1   // accumulator-drained arrive without .release.cluster (a membar per arrive; tc_common.cuh): TMEM reads are ordered by tcgen05.wait::ld + fence
#endif
constexpr int kSStages = 6;
constexpr int kSBlock = BM * BK * 2;            // 16 KB: 128 rows of one k-block

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// v[j] for a run-time j without local memory: 16 + 8 + 4 + 2 + 1 selects
__device__ __forceinline__ uint32_t mux32(const uint32_t* v, int j) {
    uint32_t a[16], b[8], c[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
    const uint32_t d0 = (j & 8) ? c[1] : c[0], d1 = (j & 8) ? c[3] : c[2];
    return (j & 16) ? d1 : d0;
}

template <bool A_RES>
__global__ void __launch_bounds__(kThreads, 1)
dedup_screen_kernel(const __grid_constant__ CUtensorMap map_hi, const ScreenParams prm) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr int kStageBytes = A_RES ? kSBlock : 2 * kSBlock;                  // B half (+ A block)
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const uint32_t a_res = base;
    const uint32_t stage0 = base + (A_RES ? (uint32_t)prm.n_kb * kSBlock : 0u);
    const uint32_t bars = stage0 + kSStages * kStageBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kSStages + s); };
    const uint32_t a_full = bars + 8u * (2 * kSStages);
    const uint32_t a_empty = a_full + 8u;
    auto tfull_bar = [&](int s) { return a_empty + 8u + 8u * s; };
    auto tempty_bar = [&](int s) { return a_empty + 8u + 16u + 8u * s; };
    const uint32_t tmem_slot = a_empty + 8u + 32u;
    const uint32_t wfull0 = tmem_slot + 8u;                 // claimed items: 8 mbarriers + 8 words, in both CTAs (gram_tc.cu, Feed)
    const uint32_t wslot0 = wfull0 + 64u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 16); }
        for (int s = 0; s < 8; ++s) mbar_init(wfull0 + 8u * s, 1);
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

    // item `seq` of this pair: every role waits for the word the leader's producer lane published (the ring of 8 cannot be
    // overrun -- the roles are within 4 items of each other, see gram_tc.cu)
    auto item_of = [&](int64_t seq) -> int64_t {
        mbar_wait_cluster(wfull0 + 8u * (uint32_t)(seq & 7), (uint32_t)((seq >> 3) & 1));
        return (int64_t)ld_shared_u32(wslot0 + 4u * (uint32_t)(seq & 7));
    };
    // item w -> unit and its tiles inside this launch's chunk (units in descending order: the long items first)
    auto unit_of = [&](int64_t w) { return prm.n_units - 1 - (prm.unit_begin + w * prm.unit_stride); };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0, a_ph = 0;
            int64_t seq_pub = 0;
            bool ended = false;
            for (int64_t seq = 0;; ++seq) {
                if (leader) {
                    // claimed when needed, not ahead: with few items per pair (50k rows: 196 items on 74 pairs) every item held
                    // in reserve is one a faster pair could have taken; the MMA warp still has the ring's stages to work on
                    while (!ended && seq_pub <= seq) {
                        const uint32_t c = atomicAdd(prm.work_counter, 1u);
                        const uint32_t slot = (uint32_t)(seq_pub & 7);
                        st_shared_u32(wslot0 + 4u * slot, c);
                        st_shared_remote_u32(wslot0 + 4u * slot, 1u, c);
                        mbar_arrive(wfull0 + 8u * slot);
                        mbar_arrive_remote<false>(wfull0 + 8u * slot, 1u);
                        ++seq_pub;
                        ended = (long long)c >= prm.n_items;
                    }
                }
                const int64_t w = item_of(seq);
                if (w >= prm.n_items) break;
                const int64_t u = unit_of(w);
                const int t1 = (int)min((int64_t)prm.t_hi, u + 1);
                const int arow = (int)(u * 256 + cta_rank * 128);
                if (A_RES) {
                    mbar_wait(a_empty, a_ph ^ 1u);
                    if (leader) mbar_expect_tx(a_full, (uint32_t)prm.n_kb * kSBlock * 2);
                    for (int kb = 0; kb < prm.n_kb; ++kb) tma_load_2d_pair(&map_hi, a_full, a_res + (uint32_t)kb * kSBlock, kb * BK, arow);
                    a_ph ^= 1u;
                }
                for (int t = prm.t_lo; t < t1; ++t) {
                    const int brow = t * 256 + (int)cta_rank * 128;        // this CTA's half of the 256 B rows
                    for (int kb = 0; kb < prm.n_kb; ++kb) {
                        mbar_wait(empty_bar(s), ph ^ 1u);
                        if (leader) mbar_expect_tx(full_bar(s), (uint32_t)kStageBytes * 2);
                        const uint32_t dst = stage0 + (uint32_t)s * kStageBytes;
                        tma_load_2d_pair(&map_hi, full_bar(s), dst, kb * BK, brow);
                        if (!A_RES) tma_load_2d_pair(&map_hi, full_bar(s), dst + kSBlock, kb * BK, arow);
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
            uint32_t ph = 0, acc_ph = 0, a_ph = 0;
            for (int64_t seq = 0;; ++seq) {
                const int64_t w = item_of(seq);
                if (w >= prm.n_items) break;
                const int64_t u = unit_of(w);
                const int t1 = (int)min((int64_t)prm.t_hi, u + 1);
                if (A_RES) { mbar_wait(a_full, a_ph); a_ph ^= 1u; }
                for (int t = prm.t_lo; t < t1; ++t) {
                    mbar_wait(tempty_bar(as), acc_ph ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)as * BN;
                    for (int kb = 0; kb < prm.n_kb; ++kb) {
                        mbar_wait(full_bar(s), ph);
                        tc_fence_after();
                        const uint32_t sb = stage0 + (uint32_t)s * kStageBytes;
                        const uint64_t b_hi = make_desc(sb);
                        const uint64_t a_hi = make_desc(A_RES ? a_res + (uint32_t)kb * kSBlock : sb + kSBlock);
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
                if (A_RES) tc_commit_pair(a_empty);          // the A block may be overwritten once these MMAs have retired
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue: thread = (row, column half)
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        int as = 0;
        uint32_t acc_ph = 0;
        for (int64_t seq = 0;; ++seq) {
            const int64_t w = item_of(seq);
            if (w >= prm.n_items) break;
            const int64_t u = unit_of(w);
            const int t1 = (int)min((int64_t)prm.t_hi, u + 1);
            const int64_t i = u * 256 + cta_rank * 128 + q * 32 + lane;     // this thread's row
            float g0 = -INFINITY, g1 = -INFINITY, g2 = -INFINITY, g3 = -INFINITY;      // the four largest Gram values over columns j < i
            uint32_t c0 = 0xffffffffu, c1 = 0xffffffffu, c2 = 0xffffffffu, c3 = 0xffffffffu;
            uint2* mine = prm.cand + (size_t)i * 8 + 4 * half;
            if (!prm.first && i < prm.N) {                   // continue the lists of the earlier chunks
                const uint2 e0 = mine[0], e1 = mine[1], e2 = mine[2], e3 = mine[3];
                g0 = __uint_as_float(e0.x); c0 = e0.y; g1 = __uint_as_float(e1.x); c1 = e1.y;
                g2 = __uint_as_float(e2.x); c2 = e2.y; g3 = __uint_as_float(e3.x); c3 = e3.y;
            }
            for (int t = prm.t_lo; t < t1; ++t) {
                mbar_wait(tfull_bar(as), acc_ph);
                tc_fence_after();
                const bool diag = ((int64_t)t == u);
                const uint32_t taddr0 = tmem_base + (uint32_t)(as * BN + half * 128) + ((uint32_t)(q * 32) << 16);
                uint32_t va[32], vb[32];
                VTF_TMEM_LD32(taddr0, va);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    // the next chunk is requested before this one is looked at (two register buffers)
                    uint32_t* v = (c & 1) ? vb : va;
                    uint32_t* vn = (c & 1) ? va : vb;
                    if (c & 1) { VTF_TMEM_WAIT(vb); } else { VTF_TMEM_WAIT(va); }
                    if (c < 3) { VTF_TMEM_LD32(taddr0 + (uint32_t)((c + 1) * 32), vn); }
                    const int64_t col0 = (int64_t)t * BN + half * 128 + c * 32;
                    if (diag) {               // only EARLIER rows (dupes.py:62): columns col0 + j >= i drop out
                        const int lim = (int)min((int64_t)32, max((int64_t)0, i - col0));
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = (j < lim) ? v[j] : 0xff800000u;      // -inf
                    }
                    float ma = fmax3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
                    float mb = fmax3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
#pragma unroll
                    for (int j = 6; j < 30; j += 4) {
                        ma = fmax3(ma, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                        mb = fmax3(mb, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    }
                    const float m = fmaxf(fmax3(ma, mb, __uint_as_float(v[30])), __uint_as_float(v[31]));
                    if (m > g3) {
                        // Rare after the first tiles: something in this chunk enters the top four.  The path is kept SMALL on
                        // purpose (a bit mask of the elements above g3, then a loop over the set bits with a 5-level select
                        // tree): fully unrolled it was 18 KB of code per chunk, the four copies thrashed the instruction cache
                        // and the kernel ran at a third of its speed (ncu: 'no_instructions' the top stall, tensor pipe 11 %).
                        uint32_t hits = 0u;
#pragma unroll
                        for (int j = 0; j < 32; ++j) hits |= (__uint_as_float(v[j]) > g3) ? (1u << j) : 0u;
                        while (hits) {             // ascending columns: an equal later column never passes an earlier one
                            const int j = __ffs((int)hits) - 1;
                            hits &= hits - 1u;
                            const float gj = __uint_as_float(mux32(v, j));
                            if (gj > g3) {         // g3 may have risen since the mask was built
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
                if (lane == 0) mbar_arrive_remote<VTF_SCREEN_RELAXED != 0>(tempty_bar(as), 0u);
                if (++as == 2) { as = 0; acc_ph ^= 1u; }
            }
            if (i < prm.N) {
                mine[0] = make_uint2(__float_as_uint(g0), c0);
                mine[1] = make_uint2(__float_as_uint(g1), c1);
                mine[2] = make_uint2(__float_as_uint(g2), c2);
                mine[3] = make_uint2(__float_as_uint(g3), c3);
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
        // Only candidates whose SCREEN value is within twice the screen's error of the best screen value can hold the exact
        // maximum (|screen - exact| <= eps * s^2 each): the others are skipped -- typically 6 of the 8 gathers of a row.
        float g_top = -INFINITY;
        for (int c = 0; c < 8; ++c) {
            const uint2 e = cand[(size_t)row * 8 + c];
            if (e.y != 0xffffffffu) g_top = fmaxf(g_top, __uint_as_float(e.x));
        }
        const float g_keep = g_top - 2.f * eps / inv_s2;
        for (int c = 0; c < 8; ++c) {
            const uint2 e = cand[(size_t)row * 8 + c];
            if ((c & 3) == 3) g_unseen = fmaxf(g_unseen, __uint_as_float(e.x));     // a half's fourth best bounds everything it did not keep
            if (e.y == 0xffffffffu || __uint_as_float(e.x) < g_keep) continue;
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

// column tiles per launch: the whole triangle when the operand fits in L2 with room to spare, else <= 48 MB of B per chunk
static int64_t screen_chunk_tiles(int64_t n_units, int d) {
    const char* e = getenv("VTF_SCREEN_CHUNK");          // read per call: the tests force short chunks on small inputs
    const int forced = e ? atoi(e) : 0;
    if (forced > 0) return min((int64_t)forced, n_units);
    const int64_t tile_bytes = 256ll * d * 2;
    if (n_units * tile_bytes <= (64ll << 20)) return n_units;
    return max((int64_t)16, (int64_t)((48ll << 20) / tile_bytes));
}

int vtf_dedup_screen_launches(int64_t N, int d) {
    if (N <= 1 || d <= 0) return 0;
    const int64_t n_units = (N + 255) / 256;
    const int64_t TC = screen_chunk_tiles(n_units, d);
    return (int)((n_units + TC - 1) / TC);
}

int vtf_dedup_screen(const void* Hi, int64_t N, int64_t P, int d, int rank, int world, void* cand, uint32_t* work_counters,
                     int n_counters, void* stream) {
    VTF_REQUIRE(Hi && cand && work_counters && N >= 0 && P >= N && P % 256 == 0 && d >= 64 && d <= 4096 && d % 64 == 0 && world > 0 &&
                    rank >= 0 && rank < world,
                "vtf_dedup_screen (P % 256 == 0, d % 64 == 0)");
    if (N <= 1) return VTF_OK;
    VTF_REQUIRE(n_counters >= vtf_dedup_screen_launches(N, d), "vtf_dedup_screen (one zeroed counter per launch: vtf_dedup_screen_launches)");
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
    const bool a_res = d <= 512;
    const size_t stage = a_res ? tc::kSBlock : 2 * tc::kSBlock;
    const size_t smem = 1024 + (a_res ? (size_t)prm.n_kb * tc::kSBlock : 0) + (size_t)tc::kSStages * stage + 384;
    if (a_res) VTF_CUDA(cudaFuncSetAttribute(tc::dedup_screen_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "vtf_dedup_screen/smem");
    else VTF_CUDA(cudaFuncSetAttribute(tc::dedup_screen_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "vtf_dedup_screen/smem");
    const int64_t TC = screen_chunk_tiles(prm.n_units, d);
    int launch_no = 0;
    for (int64_t t_lo = 0; t_lo < prm.n_units; t_lo += TC, ++launch_no) {
        prm.t_lo = (int)t_lo;
        prm.t_hi = (int)min(t_lo + TC, prm.n_units);
        prm.first = t_lo == 0;
        // this rank's units u = n_units - 1 - (rank + w * world) with u >= t_lo
        const int64_t top = prm.n_units - 1 - rank - t_lo;
        prm.n_items = top < 0 ? 0 : top / world + 1;
        prm.work_counter = work_counters + launch_no;
        if (prm.n_items <= 0) continue;
        const int grid = 2 * (int)min((int64_t)(num_sms() / 2), (int64_t)prm.n_items);
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
        if (a_res) VTF_CUDA(cudaLaunchKernelEx(&cfg, tc::dedup_screen_kernel<true>, mh, prm), "vtf_dedup_screen/launch");
        else VTF_CUDA(cudaLaunchKernelEx(&cfg, tc::dedup_screen_kernel<false>, mh, prm), "vtf_dedup_screen/launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
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
