// Tensor-core silhouette sweep for sm_100a: tcgen05.mma + TMEM accumulators + TMA operand feed.
//
// Replaces the dominant cost of the reference path: silhouette_score recomputes all N^2 euclidean
// distances for every k (grouping.py:104-105 -> SK/metrics/cluster/_unsupervised.py:310 ->
// SK/metrics/pairwise.py:2279, fp64 DGEMM on 134-row chunks) and reduces them with a Python loop of
// np.bincount (_unsupervised.py:184-189).  Here ONE pass over the Gram matrix X.X^T serves every
// labelling of the k-range.
//
// Layout (DESIGN.md "silhouette sweep"):
//   Xh   [P, d] fp16, rows permuted so that every 32-column chunk belongs to one "cell" (= one joint
//        label tuple over all labellings), zero rows as padding, values scaled by a power of two s.
//   aug  [2][P][16] fp16 (silhouette.cu, store_aug): the half squared norms m = |Xh_p|^2 / 2 as three fp16
//        pieces against a constant, so that ONE extra K=16 MMA per tile turns the Gram accumulator into
//        T = Xh_p.Xh_q - m_p - m_q = -s^2 D^2 / 2.  The norms ride in the MMA: the epilogue needs no loads.
//   cellsum [n_cells, P] fp32: per row, the sum over the columns of a cell of sqrt(max(-T, 0)) * sqrt(2)/s.
//
// Persistent, one CTA per SM, CTA PAIRS (cta_group::2) by default: the pair runs one 256 x 256 MMA tile,
// each CTA keeps its own 128 x d A block resident in shared memory (d <= 512) and streams only HALF of
// every B k-block (16 KB), 6 ring stages; both accumulator halves are double-buffered in TMEM (2 x 256
// columns) so the epilogue of tile t overlaps the MMAs of tile t+1.  Roles per CTA:
//   warp 0      TMA producer   (cp.async.bulk.tensor...cta_group::2, 128B swizzle, complete_tx on the
//                               leader's mbarrier; the augmented block travels in the same ring)
//   warp 1      MMA issuer     (leader CTA, one lane; tcgen05.mma.cta_group::2.kind::f16, M=256 N=256
//                               K=16, fp32 accumulate in TMEM; multicast tcgen05.commit frees stages and
//                               publishes accumulators in both CTAs)
//   warps 2..9  epilogue       (tcgen05.ld 32x32b.x32 -> max / sqrt.approx / add, 3 ALU ops per element,
//                               one RED.F32 per cell change; D_pp = 0 on the diagonal chunk)
// Work unit = 256-row block x column segment (segments only when row blocks cannot fill the machine).
// mbarrier waits carry a suspend-time hint (less polling = less power = higher clock under the 1 kW cap)
// and a 3 s clock64 watchdog that traps instead of hanging the device.
// Measured: 1M x 512 in 751-805 ms = 1270-1364 TFLOP/s (0.91-0.98 of the measured sustained bf16 peak);
// history of variants and same-box A/B numbers in DESIGN.md.  Algorithmic flops 2*P^2*d; tensor-bound.
#include <stdlib.h>
#include "tc_common.cuh"

namespace vtf {
namespace tc {

#ifndef VTF_EPI_WARPS
#define VTF_EPI_WARPS 8    // same-box A/B: 8 beats 16 in both forms (full 736 vs 772 ms, symmetric 407 vs 424 ms)
#endif
constexpr int kEpiWarps = VTF_EPI_WARPS;                 // 8 or 16: warps that drain TMEM (4 lane quarters x column slices)
constexpr int kSlices = kEpiWarps / 4;                   // column slices of the 256-wide tile
constexpr int kChunksPerWarp = (BN / 32) / kSlices;      // 32-column chunks per warp per tile
constexpr int kSweepThreads = 64 + 32 * kEpiWarps;
static_assert(kEpiWarps == 8 || kEpiWarps == 16, "4 lane quarters x 2 or 4 column slices");

struct Params {
    const int32_t* chunk_cell;
    float* cellsum;
    int64_t P;
    int64_t rb_begin;
    int64_t n_rb;
    int n_kb;            // d / 64
    int n_seg;           // column segments per row block
    int tiles_per_seg;   // 256-column tiles per segment
    int n_tiles;         // P / 256
    float flush_scale;   // sqrt(2) / scale: sqrt(max(-T, 0)) * flush_scale is the distance
    int sym_rank;        // SYM: this rank sweeps unit pairs q = sym_rank, sym_rank + sym_world, ...
    int sym_world;
};

// Work list shared by the three warp roles (they must walk it identically).
//   full sweep: item -> (row unit, column segment);  rbu = row unit (CPC row blocks), tiles [t0, t1)
//   SYM (D is symmetric): only tiles on or above the diagonal are computed; a tile off the diagonal feeds the row
//   sums of its rows AND, through a warp transpose-reduce, the row sums of its columns.  Unit u sweeps tiles
//   u..U-1, so units are handed out in pairs (q, U-1-q) of constant total length U+1.
template <bool PAIR, bool SYM>
__device__ __forceinline__ bool next_work(const Params& prm, int64_t it, int64_t unit0, int64_t unit_stride,
                                          int64_t& rbu, int& t0, int& t1) {
    if (SYM) {
        // Unit pair (q, U-1-q) owns the virtual tile list [0, U+1): first the U-q tiles of unit q (t = q + i), then
        // the q+1 tiles of unit U-1-q.  (Sweeping the LOWER triangle instead -- every unit starting at column tile 0,
        // which shares B tiles in L2: 109 vs 328 GB of DRAM reads -- measured 5.5% SLOWER on the same box, 449 vs
        // 425 ms: the column-sum REDs of all CTA pairs then hit the same addresses at the same time.)  Every pair has the same length, and each is cut into prm.n_seg equal
        // sub-ranges so that the work items are small against the number of CTA pairs (tail effect); a sub-range
        // that straddles the two units yields two pieces (it & 1), each with its own A block.
        const int64_t U = prm.n_rb / 2, n_q = (U + 1) / 2;
        const int64_t w = unit0 + (it >> 1) * unit_stride;             // sub-item of this rank
        const int64_t q = prm.sym_rank + (w / prm.n_seg) * prm.sym_world;
        if (q >= n_q) return false;
        const int64_t ub = U - 1 - q;
        const int64_t len_a = U - q, len_tot = (ub == q) ? len_a : U + 1;
        const int64_t L = (len_tot + prm.n_seg - 1) / prm.n_seg;
        const int64_t lo = (w % prm.n_seg) * L, hi = min(lo + L, len_tot);
        if (!(it & 1)) {                                               // piece of unit q
            rbu = q;
            t0 = (int)(q + min(lo, len_a)); t1 = (int)(q + min(hi, len_a));
        } else {                                                       // piece of unit U-1-q
            rbu = ub;
            t0 = (int)(ub + max(lo, len_a) - len_a); t1 = (int)(ub + max(hi, len_a) - len_a);
        }
        if (t1 < t0) t1 = t0;
        return true;
    }
    const int64_t n_items = (prm.n_rb / (PAIR ? 2 : 1)) * prm.n_seg;
    const int64_t item = unit0 + it * unit_stride;
    if (item >= n_items) return false;
    rbu = item / prm.n_seg;
    const int seg = (int)(item % prm.n_seg);
    t0 = seg * prm.tiles_per_seg;
    t1 = min(prm.n_tiles, t0 + prm.tiles_per_seg);
    return true;
}

// A_RESIDENT: the item's 128 x d A block is loaded once and kept; otherwise A k-blocks stream with B.
// PAIR: two CTAs of a cluster (one SM pair) run one cta_group::2 MMA of 256 rows x 256 columns: each
// CTA keeps its own 128-row A block and streams only HALF of every B tile, which halves the L2->SM
// traffic (the limiter of the single-CTA form: 64 B/cycle/SM against ~42 B/cycle/SM of L2 bandwidth)
// and the shared-memory write traffic.  CTA 0 issues the MMAs; barriers that both CTAs wait on are
// signalled with multicast commits, the accumulator-drained barrier collects arrivals from both.
template <bool A_RESIDENT, int STAGES, bool PAIR, bool SYM>
__global__ void __launch_bounds__(kSweepThreads, 1)
sil_sweep_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_aug_a, const __grid_constant__ CUtensorMap map_aug_b, const Params prm) {
    extern __shared__ uint8_t smem_raw[];
    // 128-byte swizzle needs 1024-byte aligned tiles
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr int CPC = PAIR ? 2 : 1;                        // CTAs per work unit
    constexpr int kBBytes = kBBlockBytes / CPC;              // B bytes this CTA loads per k-block
    constexpr int kStageBytes = kBBytes + (A_RESIDENT ? 0 : kABlockBytes);
    constexpr int kAugABytes = BM * 32;                      // 128 rows x 16 fp16
    constexpr int kAugBBytes = (BN / CPC) * 32;              // this CTA's B rows x 16 fp16
    static_assert(kAugABytes + kAugBBytes <= kStageBytes, "the augmented block travels in a ring slot");
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const uint32_t a_res = base;                                                  // [n_kb][16 KB] when resident
    const uint32_t stage0 = base + (A_RESIDENT ? (uint32_t)prm.n_kb * kABlockBytes : 0u);
    const uint32_t bars = stage0 + STAGES * kStageBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t a_full = bars + 8u * (2 * STAGES);
    const uint32_t a_empty = a_full + 8u;
    auto tfull_bar = [&](int s) { return a_empty + 8u + 8u * s; };
    auto tempty_bar = [&](int s) { return a_empty + 8u + 16u + 8u * s; };
    const uint32_t tmem_slot = a_empty + 8u + 32u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarps * CPC); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();   // peer barriers must exist before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    static_assert(!SYM || PAIR, "the symmetric sweep runs on CTA pairs");
    const int64_t unit0 = blockIdx.x / CPC, unit_stride = gridDim.x / CPC;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0, a_ph = 0;
            for (int64_t it = 0;; ++it) {
                int64_t rbu; int t0, t1;
                if (!next_work<PAIR, SYM>(prm, it, unit0, unit_stride, rbu, t0, t1)) break;
                if (t0 >= t1) continue;
                const int64_t rb = prm.rb_begin + CPC * rbu + cta_rank;
                // PAIR: both CTAs load; every complete_tx lands on the LEADER's barrier, which expects both halves
                if (A_RESIDENT) {
                    mbar_wait(a_empty, a_ph ^ 1u);
                    if (leader) mbar_expect_tx(a_full, (uint32_t)prm.n_kb * kABlockBytes * CPC);
                    for (int kb = 0; kb < prm.n_kb; ++kb) {
                        if (PAIR) tma_load_2d_pair(&map_a, a_full, a_res + (uint32_t)kb * kABlockBytes, kb * BK, (int)(rb * BM));
                        else tma_load_2d(&map_a, a_full, a_res + (uint32_t)kb * kABlockBytes, kb * BK, (int)(rb * BM));
                    }
                    a_ph ^= 1u;
                }
                for (int t = t0; t < t1; ++t) {
                    for (int kb = 0; kb < prm.n_kb; ++kb) {
                        mbar_wait(empty_bar(s), ph ^ 1u);
                        if (leader) mbar_expect_tx(full_bar(s), (uint32_t)kStageBytes * CPC);
                        const uint32_t dst = stage0 + (uint32_t)s * kStageBytes;
                        if (PAIR) {
                            // this CTA's half of the 256 B rows: [t*256 + rank*128, +128)
                            tma_load_2d_pair(&map_b, full_bar(s), dst, kb * BK, t * BN + (int)cta_rank * (BN / 2));
                            if (!A_RESIDENT) tma_load_2d_pair(&map_a, full_bar(s), dst + kBBytes, kb * BK, (int)(rb * BM));
                        } else {
                            tma_load_2d(&map_b, full_bar(s), dst, kb * BK, t * BN);
                            if (!A_RESIDENT) tma_load_2d(&map_a, full_bar(s), dst + kBBytes, kb * BK, (int)(rb * BM));
                        }
                        if (++s == STAGES) { s = 0; ph ^= 1u; }
                    }
                    // one more ring slot per tile: the augmented K=16 block (norm terms) of B and of A
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    if (leader) mbar_expect_tx(full_bar(s), (uint32_t)(kAugABytes + kAugBBytes) * CPC);
                    const uint32_t dst = stage0 + (uint32_t)s * kStageBytes;
                    const int bgrp = (t * BN + (int)cta_rank * (BN / CPC)) / 8;      // 8-row groups of 256 B
                    if (PAIR) {
                        tma_load_2d_pair(&map_aug_b, full_bar(s), dst, 0, bgrp);
                        tma_load_2d_pair(&map_aug_a, full_bar(s), dst + kAugBBytes, 0, (int)(rb * BM / 8));
                    } else {
                        tma_load_2d(&map_aug_b, full_bar(s), dst, 0, bgrp);
                        tma_load_2d(&map_aug_a, full_bar(s), dst + kAugBBytes, 0, (int)(rb * BM / 8));
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (lane == 0 && leader) {
            int s = 0, as = 0;
            uint32_t ph = 0, a_ph = 0, acc_ph = 0;
            for (int64_t it = 0;; ++it) {
                int64_t rbu; int t0, t1;
                if (!next_work<PAIR, SYM>(prm, it, unit0, unit_stride, rbu, t0, t1)) break;
                if (t0 >= t1) continue;
                if (A_RESIDENT) { mbar_wait(a_full, a_ph); a_ph ^= 1u; }
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(tempty_bar(as), acc_ph ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)as * BN;
                    for (int kb = 0; kb < prm.n_kb; ++kb) {
                        mbar_wait(full_bar(s), ph);
                        tc_fence_after();
                        const uint32_t b_addr = stage0 + (uint32_t)s * kStageBytes;
                        const uint32_t a_addr = A_RESIDENT ? a_res + (uint32_t)kb * kABlockBytes : b_addr + kBBytes;
                        const uint64_t a_desc = make_desc(a_addr), b_desc = make_desc(b_addr);
#pragma unroll
                        for (int k4 = 0; k4 < BK / UMMA_K; ++k4) {
                            // +32 bytes per K=16 step inside the 128-byte swizzle row: start field += 2
                            if (PAIR) tc_mma_f16_pair(d_tmem, a_desc + (uint64_t)(2 * k4), b_desc + (uint64_t)(2 * k4), kIdescPair,
                                                      (uint32_t)((kb | k4) != 0));
                            else tc_mma_f16(d_tmem, a_desc + (uint64_t)(2 * k4), b_desc + (uint64_t)(2 * k4), kIdesc,
                                            (uint32_t)((kb | k4) != 0));
                        }
                        // frees the smem stage (in both CTAs) once these MMAs retire
                        if (PAIR) tc_commit_pair(empty_bar(s)); else tc_commit(empty_bar(s));
                        if (++s == STAGES) { s = 0; ph ^= 1u; }
                    }
                    {   // T = G' - m_p - m_q: one K=16 MMA over the augmented block finishes the accumulator
                        mbar_wait(full_bar(s), ph);
                        tc_fence_after();
                        const uint32_t b_addr = stage0 + (uint32_t)s * kStageBytes;
                        const uint64_t a_desc = make_desc_aug(b_addr + kAugBBytes), b_desc = make_desc_aug(b_addr);
                        if (PAIR) tc_mma_f16_pair(d_tmem, a_desc, b_desc, kIdescPair, 1u);
                        else tc_mma_f16(d_tmem, a_desc, b_desc, kIdesc, 1u);
                        if (PAIR) tc_commit_pair(empty_bar(s)); else tc_commit(empty_bar(s));
                        if (++s == STAGES) { s = 0; ph ^= 1u; }
                    }
                    // accumulator complete -> epilogue warps of both CTAs
                    if (PAIR) tc_commit_pair(tfull_bar(as)); else tc_commit(tfull_bar(as));
                    if (++as == 2) { as = 0; acc_ph ^= 1u; }
                }
                // A block may be overwritten for the next item
                if (A_RESIDENT) { if (PAIR) tc_commit_pair(a_empty); else tc_commit(a_empty); }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (8 warps)
        const int q = warp & 3;                 // TMEM lane quarter this warp can read
        const int slice = (warp - 2) >> 2;      // which column slice of the 256-wide tile
        constexpr int kSliceCols = BN / kSlices;
        int as = 0;
        uint32_t acc_ph = 0;
        for (int64_t it = 0;; ++it) {
            int64_t rbu; int t0, t1;
            if (!next_work<PAIR, SYM>(prm, it, unit0, unit_stride, rbu, t0, t1)) break;
            if (t0 >= t1) continue;
            const int64_t rb = prm.rb_begin + CPC * rbu + cta_rank;
            const int64_t rowbase = rb * BM + q * 32;
            const int64_t p = rowbase + lane;
            const int row_cell = SYM ? __ldg(prm.chunk_cell + (rowbase >> 5)) : 0;   // the cell of this warp's 32 rows
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
            int cur_cell = -1;
            for (int t = t0; t < t1; ++t) {
                // the four cell ids of this warp's half tile, requested before the accumulator wait
#if VTF_CELL_PREFETCH
                int cells[kChunksPerWarp];
                {
                    const int32_t* cp = prm.chunk_cell + ((int64_t)t * BN + slice * kSliceCols) / 32;
                    if (kChunksPerWarp == 4) {
                        const int4 c4 = __ldg(reinterpret_cast<const int4*>(cp));
                        cells[0] = c4.x; cells[1] = c4.y; cells[kChunksPerWarp - 2] = c4.z; cells[kChunksPerWarp - 1] = c4.w;
                    } else {
                        const int2 c2 = __ldg(reinterpret_cast<const int2*>(cp));
                        cells[0] = c2.x; cells[1] = c2.y;
                    }
                }
#endif
                mbar_wait(tfull_bar(as), acc_ph);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < kChunksPerWarp; ++c) {
                    const int64_t col0 = (int64_t)t * BN + slice * kSliceCols + c * 32;
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + (uint32_t)(as * BN + slice * kSliceCols + c * 32) + ((uint32_t)(q * 32) << 16);
                    VTF_TMEM_LD32(taddr, v);
#if VTF_CELL_PREFETCH
                    const int cell = cells[c];
#else
                    const int cell = __ldg(prm.chunk_cell + (col0 >> 5));
#endif
                    if (cell != cur_cell) {
                        if (cur_cell >= 0)
                            atomicAdd(prm.cellsum + (size_t)cur_cell * prm.P + p, ((acc0 + acc1) + (acc2 + acc3)) * prm.flush_scale);
                        acc0 = acc1 = acc2 = acc3 = 0.f;
                        cur_cell = cell;
                    }
                    VTF_TMEM_WAIT(v);
                    // the accumulator holds T = s^2 (x_p.x_q - |x_p|^2/2 - |x_q|^2/2) = -s^2 D^2 / 2:
                    // distance = sqrt(max(-T, 0)) * sqrt(2)/s, the constant factor applied once per flush
                    float r[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = sqrt_approx(fmaxf(-__uint_as_float(v[j]), 0.f));
                    if (col0 == rowbase) {   // the chunk that holds the diagonal: D_pp = 0 (SK/metrics/pairwise.py:2285)
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = (j == lane) ? 0.f : r[j];
                    }
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        acc0 += r[4 * j4 + 0]; acc1 += r[4 * j4 + 1]; acc2 += r[4 * j4 + 2]; acc3 += r[4 * j4 + 3];
                    }
                    if (SYM && t != (int)rbu) {
                        // D_qp = D_pq: the same 32 x 32 block, summed over this warp's rows (all in cell row_cell), is
                        // the contribution to the row sums of columns col0..col0+31.  Transpose-reduce: after the five
                        // exchange steps lane l holds the column sum of column col0 + l.
#pragma unroll
                        for (int o = 16; o >= 1; o >>= 1) {
                            const bool up = (lane & o) != 0;
#pragma unroll
                            for (int j = 0; j < o; ++j) {
                                const float keep = up ? r[j + o] : r[j];
                                const float send = up ? r[j] : r[j + o];
                                r[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                            }
                        }
                        atomicAdd(prm.cellsum + (size_t)row_cell * prm.P + col0 + lane, r[0] * prm.flush_scale);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR) mbar_arrive_remote<SYM>(tempty_bar(as), 0u); else mbar_arrive(tempty_bar(as)); }
                if (++as == 2) { as = 0; acc_ph ^= 1u; }
            }
            if (cur_cell >= 0)
                atomicAdd(prm.cellsum + (size_t)cur_cell * prm.P + p, ((acc0 + acc1) + (acc2 + acc3)) * prm.flush_scale);
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();   // the leader's MMAs read the peer's shared memory until the end
    if (warp == 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
// aug: [P/8 groups][128 fp16 = 256 B], already in UMMA order; box = `groups` groups, no swizzle
static int make_aug_map(CUtensorMap* map, const void* aug, int64_t P, int groups) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(VTF_ERR_CUDA, "vtf_sil_sweep", "cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[2] = {128, (cuuint64_t)(P / 8)};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {128, (cuuint32_t)groups};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(aug), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[64];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(aug) failed (%d)", (int)r);
        return fail(VTF_ERR_CUDA, "vtf_sil_sweep", buf);
    }
    return VTF_OK;
}

template <bool A_RES, int STAGES, bool PAIR, bool SYM>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& maa, const CUtensorMap& mab,
                  const Params& prm, int grid, void* stream) {
    const size_t stage = kBBlockBytes / (PAIR ? 2 : 1) + (A_RES ? 0 : kABlockBytes);
    const size_t smem = 1024 + (A_RES ? (size_t)prm.n_kb * kABlockBytes : 0) + STAGES * stage + 256;
    auto kern = sil_sweep_kernel<A_RES, STAGES, PAIR, SYM>;
    VTF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "vtf_sil_sweep/smem");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kSweepThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VTF_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, maa, mab, prm), "vtf_sil_sweep/launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    VTF_CHECK_LAST("vtf_sil_sweep");
    return VTF_OK;
}

}  // namespace tc
}  // namespace vtf

using namespace vtf;

extern "C" int vtf_sil_sweep(const void* Xh, const void* aug, int64_t P, int d, float scale, const int32_t* chunk_cell,
                             int n_cells, int64_t rb_begin, int64_t rb_end, float* cellsum, void* stream) {
    VTF_REQUIRE(Xh && aug && chunk_cell && cellsum && P > 0 && P % tc::BN == 0 && d >= 64 && d <= 1024 && d % 64 == 0 &&
                    n_cells > 0 && rb_begin >= 0 && rb_end <= P / tc::BM && scale > 0.f,
                "vtf_sil_sweep (P % 256 == 0, d % 64 == 0, 64 <= d <= 1024)");
    if (rb_end <= rb_begin) return VTF_OK;
    int rc = vtf_device_check();
    if (rc != VTF_OK) return rc;
    // variant: 2 = CTA pairs (cta_group::2, default when the row-block count is even), 1 = single CTA
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("VTF_SWEEP_VARIANT"); forced = e ? atoi(e) : 0; }
    const bool even = ((rb_end - rb_begin) % 2 == 0) && (rb_begin % 2 == 0);
    const bool pair = even && forced != 1;
    CUtensorMap ma, mb, maa, mab;
    if ((rc = tc::make_operand_map(&ma, Xh, P, d, tc::BM, "vtf_sil_sweep")) != VTF_OK) return rc;
    if ((rc = tc::make_operand_map(&mb, Xh, P, d, pair ? tc::BN / 2 : tc::BN, "vtf_sil_sweep")) != VTF_OK) return rc;
    const __half* augA = (const __half*)aug;
    const __half* augB = augA + (size_t)P * 16;
    if ((rc = tc::make_aug_map(&maa, augA, P, tc::BM / 8)) != VTF_OK) return rc;
    if ((rc = tc::make_aug_map(&mab, augB, P, (pair ? tc::BN / 2 : tc::BN) / 8)) != VTF_OK) return rc;
    tc::Params prm;
    prm.chunk_cell = chunk_cell; prm.cellsum = cellsum; prm.P = P;
    prm.rb_begin = rb_begin; prm.n_rb = rb_end - rb_begin;
    prm.n_kb = d / tc::BK;
    prm.n_tiles = (int)(P / tc::BN);
    const int sms = num_sms();
    const int64_t units = pair ? prm.n_rb / 2 : prm.n_rb;
    const int64_t slots = pair ? sms / 2 : sms;
    // enough work units to fill the machine twice when the row blocks alone cannot
    int n_seg = 1;
    if (units < 2 * slots) n_seg = (int)min((int64_t)prm.n_tiles, (2 * slots + units - 1) / units);
    prm.tiles_per_seg = (prm.n_tiles + n_seg - 1) / n_seg;
    prm.n_seg = (prm.n_tiles + prm.tiles_per_seg - 1) / prm.tiles_per_seg;
    prm.flush_scale = 1.41421356237f / scale;
    prm.sym_rank = 0; prm.sym_world = 1;
    const int64_t n_units = units * prm.n_seg;
    if (pair) {
        const int grid = 2 * (int)min(slots, n_units);
        if (d <= 512) return tc::launch<true, 6, true, false>(ma, mb, maa, mab, prm, grid, stream);
        return tc::launch<false, 6, true, false>(ma, mb, maa, mab, prm, grid, stream);
    }
    const int grid = (int)min(slots, n_units);
    if (d <= 512) return tc::launch<true, 3, false, false>(ma, mb, maa, mab, prm, grid, stream);
    return tc::launch<false, 4, false, false>(ma, mb, maa, mab, prm, grid, stream);
}


// Symmetric form: D_pq = D_qp, so only tiles on or above the diagonal are computed (half the MMAs, half the L2
// traffic); every off-diagonal 32 x 32 block also feeds the row sums of its COLUMNS through a warp
// transpose-reduce.  Each rank sweeps unit pairs (q, U-1-q), q = rank, rank + world, ...; its cellsum then holds
// partial sums for ALL rows and the caller adds the ranks' arrays (all-reduce) before the finish pass.
extern "C" int vtf_sil_sweep_sym(const void* Xh, const void* aug, int64_t P, int d, float scale, const int32_t* chunk_cell,
                                 int n_cells, int rank, int world, float* cellsum, void* stream) {
    VTF_REQUIRE(Xh && aug && chunk_cell && cellsum && P > 0 && P % tc::BN == 0 && d >= 64 && d <= 1024 && d % 64 == 0 &&
                    n_cells > 0 && world > 0 && rank >= 0 && rank < world && scale > 0.f,
                "vtf_sil_sweep_sym (P % 256 == 0, d % 64 == 0, 64 <= d <= 1024)");
    int rc = vtf_device_check();
    if (rc != VTF_OK) return rc;
    CUtensorMap ma, mb, maa, mab;
    if ((rc = tc::make_operand_map(&ma, Xh, P, d, tc::BM, "vtf_sil_sweep_sym")) != VTF_OK) return rc;
    if ((rc = tc::make_operand_map(&mb, Xh, P, d, tc::BN / 2, "vtf_sil_sweep_sym")) != VTF_OK) return rc;
    const __half* augA = (const __half*)aug;
    const __half* augB = augA + (size_t)P * 16;
    if ((rc = tc::make_aug_map(&maa, augA, P, tc::BM / 8)) != VTF_OK) return rc;
    if ((rc = tc::make_aug_map(&mab, augB, P, (tc::BN / 2) / 8)) != VTF_OK) return rc;
    tc::Params prm;
    prm.chunk_cell = chunk_cell; prm.cellsum = cellsum; prm.P = P;
    prm.rb_begin = 0; prm.n_rb = P / tc::BM;
    prm.n_kb = d / tc::BK;
    prm.n_tiles = (int)(P / tc::BN);
    prm.tiles_per_seg = prm.n_tiles;
    prm.flush_scale = 1.41421356237f / scale;
    prm.sym_rank = rank; prm.sym_world = world;
    const int64_t U = prm.n_rb / 2, n_q = (U + 1) / 2;
    if (rank >= n_q) return VTF_OK;
    const int64_t mine = (n_q - rank + world - 1) / world;
    // cut every unit pair into sub-ranges until there are >= 12 work items per CTA pair (tail effect <= ~8%),
    // but keep at least 8 tiles per sub-range so the resident A block is amortised
    const int64_t slots = num_sms() / 2;
    int n_seg = 1;
    while (mine * n_seg < 12 * slots && (U + 1) / (2 * n_seg) >= 8) n_seg *= 2;
    prm.n_seg = n_seg;
    const int grid = 2 * (int)min(slots, mine * n_seg);
    if (d <= 512) return tc::launch<true, 6, true, true>(ma, mb, maa, mab, prm, grid, stream);
    return tc::launch<false, 6, true, true>(ma, mb, maa, mab, prm, grid, stream);
}
