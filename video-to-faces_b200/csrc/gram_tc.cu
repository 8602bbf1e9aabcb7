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
#ifndef VTF_EPI_FRAG
#define VTF_EPI_FRAG 1     // symmetric sweep: read TMEM in the 16x256b fragment layout (7 instead of 31 shuffles per block)
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
    int sym_rank;        // SYM, static schedule: this rank takes items w = sym_rank, sym_rank + sym_world, ...
    int sym_world;
    int chunk_tiles;     // SYM: tiles per column chunk (see next_work)
    // SYM, dynamic schedule: work items are claimed one at a time from *work_counter (atomicAdd at system scope; with
    // several ranks the counter lives in rank 0's memory and is reached over NVLink), so a slower GPU or SM pair simply
    // claims fewer items.  nullptr = the static round-robin list.  n_items = all items of ALL ranks (SYM, both forms).
    unsigned int* work_counter;
    long long n_items;
    unsigned long long* tiles_done;   // nullable: += number of 256 x 256 tiles this launch computed
    // DIAGONAL pass (full-sweep form only): every unit computes just its own diagonal tile and the epilogue stores
    // diag_out[p] = diag_scale * T_pp instead of distances.  With an all-zero augmented operand T = Xh.Xh^T, so this yields
    // the squared norms AS THE TENSOR CORE COMPUTES THEM (same instruction shape, same accumulation order as in the sweep).
    float* diag_out;
    float diag_scale;
    // debugging aid (VTF_SWEEP_TIMING=1): [0] = min over CTAs of %globaltimer at entry, [1] = max at exit
    unsigned long long* timing;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Work list shared by the three warp roles (they must walk it identically).  `w` = index of the work item.
//   full sweep: item -> (row unit, column segment);  rbu = row unit (CPC row blocks), tiles [t0, t1)
//   SYM (D is symmetric): only tiles on or above the diagonal are computed; a tile off the diagonal feeds the row
//   sums of its rows AND, through a warp transpose-reduce, the row sums of its columns.  The triangle is cut into
//   column CHUNKS of prm.chunk_tiles tiles (<= 32 MB of B operand, a quarter of the 126 MB L2) and walked chunk by
//   chunk: item (c, u) = unit u against the tiles of chunk c at or right of the diagonal, items ordered c-major.
//   CTA pairs that run at the same time therefore read the SAME chunk of B -- it comes from DRAM once and is then
//   served by L2 for every unit -- and consecutive units start at different offsets inside the chunk (`o`, walked
//   with wrap-around), so their column-sum REDs never hit the same addresses at the same time.  (History, same-box
//   A/Bs: unit pairs (q, U-1-q) swept end to end, 330 GB of DRAM reads, 409-451 ms; the same cut into 4 sub-ranges
//   and claimed dynamically, 509 ms -- four column regions in flight thrash L2; the lower triangle with every unit
//   starting at column 0, 109 GB but 449 vs 425 ms because all pairs RED into the same columns at once.)
template <bool PAIR, bool SYM>
__device__ __forceinline__ bool next_work(const Params& prm, int64_t w, int64_t& rbu, int& t0, int& t1, int& o) {
    o = 0;
    if (SYM) {
        if (w >= prm.n_items) return false;
        const int64_t U = prm.n_rb / 2, TC = prm.chunk_tiles;
        const int64_t nc_full = U / TC;                                 // chunks whose unit count is (c+1) * TC
        int64_t c = (int64_t)((sqrt(8.0 * (double)w / (double)TC + 1.0) - 1.0) * 0.5);
        c = min(max(c, (int64_t)0), nc_full);
        while (c < nc_full && TC * (c + 1) * (c + 2) / 2 <= w) ++c;     // base(c+1) <= w
        while (c > 0 && TC * c * (c + 1) / 2 > w) --c;                  // base(c) > w
        const int64_t u = w - TC * c * (c + 1) / 2;
        const int64_t lo = max(c * TC, u), hi = min((c + 1) * TC, U);
        rbu = u;
        t0 = (int)lo; t1 = (int)hi;
        o = (int)(u % (hi - lo));
        return true;
    }
    const int64_t n_items = (prm.n_rb / (PAIR ? 2 : 1)) * prm.n_seg;
    const int64_t item = w;
    if (item >= n_items) return false;
    rbu = item / prm.n_seg;
    const int seg = (int)(item % prm.n_seg);
    t0 = seg * prm.tiles_per_seg;
    t1 = min(prm.n_tiles, t0 + prm.tiles_per_seg);
    if (prm.diag_out) { t0 = (int)((prm.rb_begin + (PAIR ? 2 : 1) * rbu) * BM / BN); t1 = t0 + 1; }   // the tile that holds the unit's diagonal
    return true;
}

// Where the warp roles get their work item from.  Static: item = unit0 + seq * stride, computed by every role on its
// own.  Dynamic (SYM with a work counter): lane 0 of the leader CTA's producer warp claims items with an atomicAdd and
// publishes each one -- a word in an 8-deep ring in the shared memory of BOTH CTAs of the pair plus an mbarrier
// arrive -- always one item ahead of its own use; every other role waits on that barrier and reads the word.  The
// ring cannot be overrun: the producer is at most STAGES k-blocks ahead of the MMA warp, the MMA warp at most two
// tiles ahead of the slowest epilogue warp, and every item has at least one tile, so all roles are within 4 items of
// each other.  The first claimed index >= n_items is published too and ends every role's loop.
struct Feed {
    bool dyn;
    uint32_t wfull0, wslot0;
    int64_t unit0, stride;
    int rank, world;                     // static list: ranks interleave item by item
    __device__ __forceinline__ int64_t get(int64_t seq) const {
        if (!dyn) return (unit0 + seq * stride) * world + rank;
        mbar_wait_cluster(wfull0 + 8u * (uint32_t)(seq & 7), (uint32_t)((seq >> 3) & 1));
        return (int64_t)ld_shared_u32(wslot0 + 4u * (uint32_t)(seq & 7));
    }
};

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
    const uint32_t wfull0 = tmem_slot + 8u;                  // dynamic schedule: 8 mbarriers + 8 words (see Feed)
    const uint32_t wslot0 = wfull0 + 64u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (prm.timing && threadIdx.x == 0) atomicMin(prm.timing, globaltimer_ns());

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarps * CPC); }
        for (int s = 0; s < 8; ++s) mbar_init(wfull0 + 8u * s, 1);
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
    Feed feed;
    feed.dyn = SYM && prm.work_counter != nullptr;
    feed.wfull0 = wfull0; feed.wslot0 = wslot0;
    feed.unit0 = blockIdx.x / CPC; feed.stride = gridDim.x / CPC;
    feed.rank = SYM ? prm.sym_rank : 0; feed.world = SYM ? prm.sym_world : 1;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0, a_ph = 0;
            int64_t w = 0, seq_pub = 0;
            bool ended = false;
            for (int64_t it = 0;; ++it) {
                int64_t rbu; int t0, t1, o;
                {
                    const int64_t seq = it;
                    if (feed.dyn && leader) {
                        // claim and publish up to one item ahead of the one about to be loaded
                        while (!ended && seq_pub <= seq + 1) {
                            const uint32_t c = atomicAdd_system(prm.work_counter, 1u);
                            const uint32_t slot = (uint32_t)(seq_pub & 7);
                            st_shared_u32(wslot0 + 4u * slot, c);
                            st_shared_remote_u32(wslot0 + 4u * slot, 1u, c);
                            mbar_arrive(wfull0 + 8u * slot);
                            mbar_arrive_remote<false>(wfull0 + 8u * slot, 1u);
                            ++seq_pub;
                            ended = (long long)c >= prm.n_items;
                        }
                    }
                    w = feed.get(seq);
                }
                if (!next_work<PAIR, SYM>(prm, w, rbu, t0, t1, o)) break;
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
                for (int i = 0; i < t1 - t0; ++i) {
                    const int t = t0 + ((o + i >= t1 - t0) ? o + i - (t1 - t0) : o + i);     // start at offset o, wrap
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
            int64_t w = 0;
            unsigned long long n_done = 0;
            for (int64_t it = 0;; ++it) {
                int64_t rbu; int t0, t1, o;
                w = feed.get(it);
                if (!next_work<PAIR, SYM>(prm, w, rbu, t0, t1, o)) break;
                if (t0 >= t1) continue;
                n_done += (unsigned long long)(t1 - t0);
                if (A_RESIDENT) { mbar_wait(a_full, a_ph); a_ph ^= 1u; }
                for (int i = 0; i < t1 - t0; ++i) {
                    const int t = t0 + ((o + i >= t1 - t0) ? o + i - (t1 - t0) : o + i);
                    (void)t;
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
            if (prm.tiles_done && n_done) atomicAdd(prm.tiles_done, n_done);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (8 warps)
        const int q = warp & 3;                 // TMEM lane quarter this warp can read
        const int slice = (warp - 2) >> 2;      // which column slice of the 256-wide tile
        constexpr int kSliceCols = BN / kSlices;
        constexpr bool FRAG = SYM && (VTF_EPI_FRAG != 0);
        const int g = lane >> 2, tq = lane & 3;
        const bool diag = !SYM && prm.diag_out != nullptr;   // diagonal pass: store T_pp, no cells, no distances
        int as = 0;
        uint32_t acc_ph = 0;
        int64_t w = 0;
        for (int64_t it = 0;; ++it) {
            int64_t rbu; int t0, t1, o;
            w = feed.get(it);
            if (!next_work<PAIR, SYM>(prm, w, rbu, t0, t1, o)) break;
            if (t0 >= t1) continue;
            const int64_t rb = prm.rb_begin + CPC * rbu + cta_rank;
            const int64_t rowbase = rb * BM + q * 32;
            const int64_t p = rowbase + lane;
            const int row_cell = SYM ? __ldg(prm.chunk_cell + (rowbase >> 5)) : 0;   // the cell of this warp's 32 rows
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
            float2 fa0 = make_float2(0.f, 0.f), fa1 = fa0, fa2 = fa0, fa3 = fa0;      // fragment layout: (even, odd) columns
            int cur_cell = -1;
            // row sums of the finished cell run -> cellsum (warp-uniform call).  One row per thread in the 32x32b layout;
            // in the fragment layout the four threads of a group first add their column partials (2 shuffles per
            // accumulator) and thread tq then owns row g + 8 tq.
            auto flush_rows = [&](int cell_id) {
                if (FRAG) {
                    acc0 = fa0.x + fa0.y; acc1 = fa1.x + fa1.y; acc2 = fa2.x + fa2.y; acc3 = fa3.x + fa3.y;
                    acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1); acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
                    acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
                    acc2 += __shfl_xor_sync(0xffffffffu, acc2, 1); acc2 += __shfl_xor_sync(0xffffffffu, acc2, 2);
                    acc3 += __shfl_xor_sync(0xffffffffu, acc3, 1); acc3 += __shfl_xor_sync(0xffffffffu, acc3, 2);
                    const float mine = tq == 0 ? acc0 : tq == 1 ? acc1 : tq == 2 ? acc2 : acc3;
                    atomicAdd(prm.cellsum + (size_t)cell_id * prm.P + rowbase + g + 8 * tq, mine * prm.flush_scale);
                } else {
                    atomicAdd(prm.cellsum + (size_t)cell_id * prm.P + p, ((acc0 + acc1) + (acc2 + acc3)) * prm.flush_scale);
                }
            };
            for (int i = 0; i < t1 - t0; ++i) {
                const int t = t0 + ((o + i >= t1 - t0) ? o + i - (t1 - t0) : o + i);
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
                    if (FRAG) {
                        VTF_TMEM_LD16X256(taddr, v, 0);                       // rows  0..15 of this warp's 32
                        VTF_TMEM_LD16X256(taddr + (16u << 16), v, 16);        // rows 16..31
                    } else {
                        VTF_TMEM_LD32(taddr, v);
                    }
#if VTF_CELL_PREFETCH
                    const int cell = diag ? 0 : cells[c];
#else
                    const int cell = diag ? 0 : __ldg(prm.chunk_cell + (col0 >> 5));
#endif
                    if (cell != cur_cell) {
                        if (cur_cell >= 0) flush_rows(cur_cell);
                        acc0 = acc1 = acc2 = acc3 = 0.f;
                        fa0 = fa1 = fa2 = fa3 = make_float2(0.f, 0.f);
                        cur_cell = cell;
                    }
                    VTF_TMEM_WAIT(v);
                    // the accumulator holds T = s^2 (x_p.x_q - |x_p|^2/2 - |x_q|^2/2) = -s^2 D^2 / 2:
                    // distance = sqrt(max(-T, 0)) * sqrt(2)/s, the constant factor applied once per flush
                    float r[32];
                    if (FRAG) {
                        // |T| instead of max(-T, 0): the two differ only where rounding noise made the squared distance of
                        // (nearly) coincident rows negative, where either answer is noise of the same size -- and the
                        // absolute value rides on the MUFU operand for free (one FMNMX less per element)
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = sqrt_approx(fabsf(__uint_as_float(v[j])));
                        // register j: row 16 (j / 16) + 8 ((j / 2) % 2) + g, column 8 ((j % 16) / 4) + 2 tq + j % 2
                        if (col0 == rowbase) {   // the chunk that holds the diagonal: D_pp = 0 (SK/metrics/pairwise.py:2285)
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int row = 16 * (j >> 4) + 8 * ((j >> 1) & 1) + g, col = 8 * ((j & 15) >> 2) + 2 * tq + (j & 1);
                                r[j] = (row == col) ? 0.f : r[j];
                            }
                        }
                        // row sums, two columns at a time (packed fp32x2 adds): fa0..3 = rows g, g + 8, g + 16, g + 24 (this
                        // thread's 8 columns; the other three threads of the group hold the rest -- combined at flush time)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            fa0 = __fadd2_rn(fa0, make_float2(r[4 * i + 0], r[4 * i + 1]));
                            fa1 = __fadd2_rn(fa1, make_float2(r[4 * i + 2], r[4 * i + 3]));
                            fa2 = __fadd2_rn(fa2, make_float2(r[16 + 4 * i + 0], r[16 + 4 * i + 1]));
                            fa3 = __fadd2_rn(fa3, make_float2(r[16 + 4 * i + 2], r[16 + 4 * i + 3]));
                        }
                        if (t != (int)rbu) {
                            // D_qp = D_pq: the block summed over this warp's 32 rows (all in cell row_cell) is the contribution
                            // to the row sums of columns col0..col0+31.  Four rows are added inside the thread; the eight
                            // threads that share tq then transpose-reduce their 8 column partials in 4 + 2 + 1 shuffles, after
                            // which the thread holds column 8 (g / 2) + 2 tq + g % 2.
                            float cs[8];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 c2 = __fadd2_rn(
                                    __fadd2_rn(make_float2(r[4 * i + 0], r[4 * i + 1]), make_float2(r[4 * i + 2], r[4 * i + 3])),
                                    __fadd2_rn(make_float2(r[16 + 4 * i + 0], r[16 + 4 * i + 1]), make_float2(r[16 + 4 * i + 2], r[16 + 4 * i + 3])));
                                cs[2 * i + 0] = c2.x;
                                cs[2 * i + 1] = c2.y;
                            }
#pragma unroll
                            for (int o = 4; o >= 1; o >>= 1) {                 // lane bit 4 o: 16, 8, 4
                                const bool up = (lane & (4 * o)) != 0;
#pragma unroll
                                for (int j = 0; j < o; ++j) {
                                    const float keep = up ? cs[j + o] : cs[j];
                                    const float send = up ? cs[j] : cs[j + o];
                                    cs[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4 * o);
                                }
                            }
                            atomicAdd(prm.cellsum + (size_t)row_cell * prm.P + col0 + 8 * (g >> 1) + 2 * tq + (g & 1),
                                      cs[0] * prm.flush_scale);
                        }
                    } else {
                        if (diag) {                              // diagonal pass: store T_pp, no distances (warp-uniform branch)
                            if (col0 == rowbase) {
                                float dv = 0.f;
#pragma unroll
                                for (int j = 0; j < 32; ++j) dv = (j == lane) ? __uint_as_float(v[j]) : dv;
                                prm.diag_out[p] = dv * prm.diag_scale;
                            }
                            continue;
                        }
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
                            // one TMEM lane = one row per thread: transpose-reduce over all 32 lanes, 16 + 8 + 4 + 2 + 1 shuffles;
                            // afterwards lane l holds the column sum of column col0 + l
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
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR) mbar_arrive_remote<SYM>(tempty_bar(as), 0u); else mbar_arrive(tempty_bar(as)); }
                if (++as == 2) { as = 0; acc_ph ^= 1u; }
            }
            if (cur_cell >= 0 && !diag) flush_rows(cur_cell);
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();   // the leader's MMAs read the peer's shared memory until the end
    if (prm.timing && threadIdx.x == 0) atomicMax(prm.timing + 1, globaltimer_ns());
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
    const size_t smem = 1024 + (A_RES ? (size_t)prm.n_kb * kABlockBytes : 0) + STAGES * stage + 384;
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
    prm.work_counter = nullptr; prm.n_items = 0; prm.tiles_done = nullptr; prm.chunk_tiles = 0;
    prm.diag_out = nullptr; prm.diag_scale = 0.f; prm.timing = nullptr;
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


// Diagonal pass: out[p] = out_scale * T_pp for every row, T as the sweeps compute it (CTA pairs, M = 256, N = 256, K = 16,
// k-blocks in order, the augmented block last).  With an all-zero `aug`, T_pp = (Xh.Xh^T)_pp: the squared norm of row p in
// the TENSOR CORE's arithmetic -- fp32 accumulation with truncation after every MMA, ~3e-6 low after 32 steps.  Norms taken
// from here make T_pq = g_pq - g_pp/2 - g_qq/2 cancel bit for bit when rows p and q coincide and remove the common bias for
// rows that nearly do; with fp32-FMA norms (vtf_sil_pack) that bias is a floor of ~2e-3 |x| under every small distance.
// P/256 tiles = P * 256 * 2 * d flop: 0.03 % of a sweep.
extern "C" int vtf_sil_diag(const void* Xh, const void* aug, int64_t P, int d, float out_scale, float* out, void* stream) {
    VTF_REQUIRE(Xh && aug && out && P > 0 && P % tc::BN == 0 && d >= 64 && d <= 1024 && d % 64 == 0, "vtf_sil_diag");
    int rc = vtf_device_check();
    if (rc != VTF_OK) return rc;
    CUtensorMap ma, mb, maa, mab;
    if ((rc = tc::make_operand_map(&ma, Xh, P, d, tc::BM, "vtf_sil_diag")) != VTF_OK) return rc;
    if ((rc = tc::make_operand_map(&mb, Xh, P, d, tc::BN / 2, "vtf_sil_diag")) != VTF_OK) return rc;
    const __half* augA = (const __half*)aug;
    const __half* augB = augA + (size_t)P * 16;
    if ((rc = tc::make_aug_map(&maa, augA, P, tc::BM / 8)) != VTF_OK) return rc;
    if ((rc = tc::make_aug_map(&mab, augB, P, (tc::BN / 2) / 8)) != VTF_OK) return rc;
    tc::Params prm;
    prm.chunk_cell = (const int32_t*)out;           // not read in this mode (the cell look-up of the epilogue needs a valid address)
    prm.cellsum = out; prm.P = P;
    prm.rb_begin = 0; prm.n_rb = P / tc::BM;
    prm.n_kb = d / tc::BK;
    prm.n_tiles = (int)(P / tc::BN);
    prm.n_seg = 1; prm.tiles_per_seg = prm.n_tiles;
    prm.flush_scale = 0.f;
    prm.sym_rank = 0; prm.sym_world = 1;
    prm.work_counter = nullptr; prm.n_items = 0; prm.tiles_done = nullptr; prm.chunk_tiles = 0;
    prm.diag_out = out; prm.diag_scale = out_scale; prm.timing = nullptr;
    const int64_t units = prm.n_rb / 2;
    const int grid = 2 * (int)min((int64_t)(num_sms() / 2), units);
    if (d <= 512) return tc::launch<true, 6, true, false>(ma, mb, maa, mab, prm, grid, stream);
    return tc::launch<false, 6, true, false>(ma, mb, maa, mab, prm, grid, stream);
}

// VTF_SWEEP_TIMING=1: the symmetric sweep records its own first-entry / last-exit %globaltimer values (a cross-check of the
// event-based timings); vtf_debug_sweep_timing returns the last launch's pair.
static unsigned long long* g_sweep_timing = nullptr;
static unsigned long long* sweep_timing_buffer(void* stream) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("VTF_SWEEP_TIMING"); on = (e && atoi(e)) ? 1 : 0; }
    if (!on) return nullptr;
    if (!g_sweep_timing && cudaMalloc(&g_sweep_timing, 16) != cudaSuccess) { g_sweep_timing = nullptr; return nullptr; }
    const unsigned long long init[2] = {~0ull, 0ull};
    cudaMemcpyAsync(g_sweep_timing, init, 16, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    return g_sweep_timing;
}

extern "C" int vtf_debug_sweep_timing(uint64_t* out2) {
    if (!g_sweep_timing || !out2) return VTF_ERR_ARG;
    return cudaMemcpy(out2, g_sweep_timing, 16, cudaMemcpyDeviceToHost) == cudaSuccess ? VTF_OK : VTF_ERR_CUDA;
}

// Symmetric form: D_pq = D_qp, so only tiles on or above the diagonal are computed (half the MMAs, half the L2
// traffic); every off-diagonal 32 x 32 block also feeds the row sums of its COLUMNS through a warp
// transpose-reduce.  Each rank sweeps unit pairs (q, U-1-q), q = rank, rank + world, ...; its cellsum then holds
// partial sums for ALL rows and the caller adds the ranks' arrays (all-reduce) before the finish pass.
// With `work_counter` (zero before the first rank starts; shared by all ranks, see vtf_shared_alloc) the unit pairs are
// not dealt out by rank: every CTA pair of every rank claims the next sub-range from the counter.
extern "C" int vtf_sil_sweep_sym(const void* Xh, const void* aug, int64_t P, int d, float scale, const int32_t* chunk_cell,
                                 int n_cells, int rank, int world, float* cellsum, uint32_t* work_counter,
                                 uint64_t* tiles_done, void* stream) {
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
    prm.work_counter = work_counter; prm.n_items = 0; prm.tiles_done = (unsigned long long*)tiles_done;
    prm.diag_out = nullptr; prm.diag_scale = 0.f;
    prm.timing = sweep_timing_buffer(stream);
    const int64_t U = prm.n_rb / 2;
    // column chunk: <= 32 MB of B operand (a quarter of the 126 MB L2, so the chunk being left and the one being
    // entered both stay resident), 16..256 tiles; VTF_SWEEP_CHUNK overrides.  Same-box A/B at 1M x 512 (static /
    // claimed): 32 tiles 426 / 422 ms, 64 tiles 418 / 414 ms, 128 tiles 415 / 409 ms.
    static int forced_chunk = -1;
    if (forced_chunk < 0) { const char* e = getenv("VTF_SWEEP_CHUNK"); forced_chunk = e ? atoi(e) : 0; }
    int64_t TC = (32ll << 20) / (256ll * d * 2);
    TC = max((int64_t)16, min((int64_t)256, TC));
    if (forced_chunk > 0) TC = forced_chunk;
    TC = min(TC, U);
    prm.chunk_tiles = (int)TC;
    prm.n_seg = 1;
    const int64_t nc_full = U / TC;
    prm.n_items = TC * nc_full * (nc_full + 1) / 2 + (U % TC ? U : 0);
    const int64_t slots = num_sms() / 2;
    if (work_counter) { prm.sym_rank = 0; prm.sym_world = 1; }        // claimed, not dealt out
    const int64_t mine = work_counter ? prm.n_items : (prm.n_items - rank + world - 1) / world;
    if (mine <= 0) return VTF_OK;
    const int grid = 2 * (int)min(slots, mine);
    if (d <= 512) return tc::launch<true, 6, true, true>(ma, mb, maa, mab, prm, grid, stream);
    return tc::launch<false, 6, true, true>(ma, mb, maa, mab, prm, grid, stream);
}
