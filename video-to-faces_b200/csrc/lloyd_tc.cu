// "Wide" Lloyd iterations for sm_100a: the E-step of a whole k-range as ONE dense contraction X.C^T on tcgen05.
//
// grouping.py:99-101 fits KMeans once per k; with --clusters '2-32' (BASELINE config 4: 1M x 1024, 527 centres in
// total, 54-100 iterations for the larger k) the E-step is 2*N*sum(k)*d = 1.1e12 flop per lock-step iteration --
// a GEMM, not a stream: 43 ms on the fp32 pipe (rowtile.cu / rowdot.cu), ~4 ms here.
//
//   wide_assign_kernel   columns = the centres of every ACTIVE labelling, concatenated; CTA pairs (cta_group::2) own
//                        256 rows and walk the column tiles; operands split into two fp16 pieces (s*x = hi + lo,
//                        hi.hi + hi.lo + lo.hi, fp32 accumulate in TMEM, as gram_tc_dedup.cu) so that the score
//                        |c|^2/2 - x.c carries fp32-GEMM accuracy.  Epilogue: one thread = one row; running
//                        argmin per labelling (first index on ties, _k_means_lloyd.pyx:205-213) plus the SET of
//                        centres whose score lies within the error bound of the best one.  Rows with more than one
//                        candidate are flagged (sign bit of the label, candidate mask beside it) ...
//   wide_refine_kernel   ... and are decided by exact fp32 dot products in rowdot.cu's summation order, so the labels
//                        equal those of the fp32 kernels.
//   wide_count / commit  cluster sizes, "labels changed" counts (strict convergence, _kmeans.py:723) and the list of
//                        (row, old, new) changes of every labelling.
//   wide_sums_delta/full M-step sums in 64-bit FIXED POINT (x * 2^S rounded once): integer adds are exact and
//                        order-independent, so sums updated by +x / -x for the rows that changed are bit-identical
//                        to sums recomputed from scratch (and to an all-reduce over ranks in any order).  After the
//                        first iterations < 1 % of the rows move per iteration: the M-step costs those rows only.
//   wide_update_kernel   C_new = sums / counts, centre shift, per-labelling stats; one CTA per labelling.
#include "tc_common.cuh"

namespace vtf {
namespace tc {

constexpr int kWThreads = 192;                  // warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue
constexpr int kWStages = 3;
constexpr int kWBlock = BM * BK * 2;            // 16 KB: 128 rows x 64 fp16
constexpr int kWStageBytes = 4 * kWBlock;       // B_hi, B_lo, A_hi, A_lo
constexpr int kWMaxCols = 2048;                 // column table held in shared memory (8 tiles of 256)
constexpr uint32_t kColPad = 1u, kColLast = 2u; // info word: flags | local index << 4 (16 bits) | labelling << 20

struct WideParams {
    const float2* tab;            // [Kp] {|c|^2 / 2 (+inf for padding), info word as float bits}
    const float* rown;            // [P] |x| per row (0 for padding rows)
    const unsigned int* cmax2h;   // float bits of max_c |c|^2 / 2
    int32_t* labels_new;          // [n_lab][label_stride]
    long long label_stride;
    long long n_rows;
    long long n_units;            // P / 256
    int n_tiles;                  // column tiles of 256; the last one is n_last wide
    int n_last;                   // columns of the last tile, a multiple of 32: its MMAs and epilogue are that narrow
    int n_kb;                     // d / 64
    float inv;                    // 1 / scale^2
    float eps;                    // relative error bound of the tensor-core score, times |x| * max|c|
};

__global__ void __launch_bounds__(kWThreads, 1)
wide_assign_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                   const __grid_constant__ CUtensorMap map_ch, const __grid_constant__ CUtensorMap map_cl,
                   const __grid_constant__ CUtensorMap map_ch_last, const __grid_constant__ CUtensorMap map_cl_last,
                   const WideParams prm) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const uint32_t stage0 = base;
    const uint32_t bars = stage0 + kWStages * kWStageBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kWStages + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * kWStages) + 8u * s; };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * kWStages) + 16u + 8u * s; };
    const uint32_t tmem_slot = bars + 8u * (2 * kWStages) + 32u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float2* s_tab = reinterpret_cast<float2*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < prm.n_tiles * BN; i += kWThreads) s_tab[i] = prm.tab[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kWStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8); }   // 4 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const long long pair0 = blockIdx.x / 2, n_pairs = gridDim.x / 2;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long u = pair0; u < prm.n_units; u += n_pairs) {
                const int arow = (int)(u * 256 + cta_rank * 128);
                for (int t = 0; t < prm.n_tiles; ++t) {
                    // this CTA's half of the tile's centre rows; a narrow last tile has its own maps whose box is n_last / 2 rows,
                    // so a k-range with few centres (k = 2..6: 32 columns) streams the row operand and little else
                    const bool last = t == prm.n_tiles - 1;
                    const int half = (last ? prm.n_last : BN) / 2;
                    const int brow = t * 256 + (int)cta_rank * half;
                    const uint32_t tx = 2u * (uint32_t)(2 * kWBlock + 2 * half * BK * 2);       // both CTAs: A hi/lo + B hi/lo
                    const CUtensorMap* mh = last ? &map_ch_last : &map_ch;
                    const CUtensorMap* ml = last ? &map_cl_last : &map_cl;
                    for (int kb = 0; kb < prm.n_kb; ++kb) {
                        mbar_wait(empty_bar(s), ph ^ 1u);
                        if (leader) mbar_expect_tx(full_bar(s), tx);
                        const uint32_t dst = stage0 + (uint32_t)s * kWStageBytes;
                        tma_load_2d_pair(mh, full_bar(s), dst, kb * BK, brow);
                        tma_load_2d_pair(ml, full_bar(s), dst + kWBlock, kb * BK, brow);
                        tma_load_2d_pair(&map_xh, full_bar(s), dst + 2 * kWBlock, kb * BK, arow);
                        tma_load_2d_pair(&map_xl, full_bar(s), dst + 3 * kWBlock, kb * BK, arow);
                        if (++s == kWStages) { s = 0; ph ^= 1u; }
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
            for (long long u = pair0; u < prm.n_units; u += n_pairs) {
                for (int t = 0; t < prm.n_tiles; ++t) {
                    mbar_wait(tempty_bar(as), acc_ph ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)as * BN;
                    const uint32_t n_t = (uint32_t)(t == prm.n_tiles - 1 ? prm.n_last : BN);
                    const uint32_t idesc = (1u << 4) | ((n_t >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
                    for (int kb = 0; kb < prm.n_kb; ++kb) {
                        mbar_wait(full_bar(s), ph);
                        tc_fence_after();
                        const uint32_t sb = stage0 + (uint32_t)s * kWStageBytes;
                        const uint64_t b_hi = make_desc(sb), b_lo = make_desc(sb + kWBlock);
                        const uint64_t a_hi = make_desc(sb + 2 * kWBlock), a_lo = make_desc(sb + 3 * kWBlock);
#pragma unroll
                        for (int k4 = 0; k4 < BK / UMMA_K; ++k4) {
                            const uint64_t o = (uint64_t)(2 * k4);
                            tc_mma_f16_pair(d_tmem, a_hi + o, b_hi + o, idesc, (uint32_t)((kb | k4) != 0));
                            tc_mma_f16_pair(d_tmem, a_hi + o, b_lo + o, idesc, 1u);
                            tc_mma_f16_pair(d_tmem, a_lo + o, b_hi + o, idesc, 1u);
                        }
                        tc_commit_pair(empty_bar(s));
                        if (++s == kWStages) { s = 0; ph ^= 1u; }
                    }
                    tc_commit_pair(tfull_bar(as));
                    if (++as == 2) { as = 0; acc_ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue: thread = row, all 256 columns of a tile
        // One warp per scheduler, so the loop must be straight-line code with instruction-level parallelism: no branch per
        // element (the stores at the end of a labelling are predicated PTX), best +- tol kept as running values so that the
        // only loop-carried chain is compare -> select, and the TMEM load of the next 32 columns in flight behind the current.
        const int q = warp & 3;                      // TMEM lane quarter this warp may read
        int as = 0;
        uint32_t acc_ph = 0;
        const float cmax = sqrtf(2.f * __uint_as_float(*prm.cmax2h));
        for (long long u = pair0; u < prm.n_units; u += n_pairs) {
            const long long row = u * 256 + cta_rank * 128 + q * 32 + lane;
            const float tol = prm.eps * prm.rown[row] * cmax;
            const uint32_t live = row < prm.n_rows ? 1u : 0u;
            const uint32_t stride4 = (uint32_t)prm.label_stride * 4u, row4 = (uint32_t)row * 4u;
            char* lab_base = reinterpret_cast<char*>(prm.labels_new);
            float best = INFINITY, lo = INFINITY, hi = INFINITY;     // lo = best - tol, hi = best + tol
            uint32_t arg = 0u, mask = 0u, multi = 0u;
            auto process = [&](const uint32_t (&v)[32], const float2* tb) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float2 e = tb[j];
                    const uint32_t info = __float_as_uint(e.y);
                    const float s = fmaf(-__uint_as_float(v[j]), prm.inv, e.x);      // |c|^2/2 - x.c
                    const float s_lo = s - tol, s_hi = s + tol;
                    const uint32_t local = (info >> 4) & 0xffffu;
                    const uint32_t bit = 0x100u << (local & 15u);     // candidate bits sit at [8, 24) of the label word
                    const bool p_clear = s < lo;             // beats everything seen by more than the error bound
                    const bool p_cand = s <= hi;             // within the bound of the best so far (indices 16 apart share a bit)
                    const bool p_new = s < best;             // strict: the first index wins ties
                    mask = p_clear ? bit : (p_cand ? (mask | bit) : mask);
                    multi = p_clear ? 0u : (multi | (p_cand ? 0x80000000u : 0u));
                    best = p_new ? s : best; lo = p_new ? s_lo : lo; hi = p_new ? s_hi : hi; arg = p_new ? local : arg;
                    const uint32_t last = (info >> 1) & 1u;  // warp-uniform: last centre of this labelling
                    // label word = index | candidate bits | flag; one predicated store, byte offset in 32 bits
                    // (n_lab * label_stride * 4 < 2^32, checked on the host)
                    const uint32_t ob = (info >> 20) * stride4 + row4;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t.reg .u64 a;\n\t"
                        "setp.ne.u32 p, %0, 0;\n\t"
                        "cvt.u64.u32 a, %1;\n\t"
                        "add.u64 a, a, %2;\n\t"
                        "@p st.global.u32 [a], %3;\n\t}"
                        ::"r"(last & live), "r"(ob), "l"(lab_base), "r"(arg | multi | (multi ? mask : 0u)));
                    best = last ? INFINITY : best; lo = last ? INFINITY : lo; hi = last ? INFINITY : hi;
                    mask = last ? 0u : mask; multi = last ? 0u : multi;
                }
            };
            for (int t = 0; t < prm.n_tiles; ++t) {
                mbar_wait(tfull_bar(as), acc_ph);
                tc_fence_after();
                const int n_chunks = (t == prm.n_tiles - 1 ? prm.n_last : BN) / 32;
                const uint32_t taddr = tmem_base + (uint32_t)(as * BN) + ((uint32_t)(q * 32) << 16);
                const float2* tb = s_tab + t * BN;
                uint32_t va[32], vb[32];
                VTF_TMEM_LD32(taddr, va);
#pragma unroll 1
                for (int c = 0; c < n_chunks; c += 2) {
                    VTF_TMEM_WAIT(va);
                    if (c + 1 < n_chunks) VTF_TMEM_LD32(taddr + (uint32_t)((c + 1) * 32), vb);
                    process(va, tb + c * 32);
                    if (c + 1 < n_chunks) {
                        VTF_TMEM_WAIT(vb);
                        if (c + 2 < n_chunks) VTF_TMEM_LD32(taddr + (uint32_t)((c + 2) * 32), va);
                        process(vb, tb + (c + 1) * 32);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote<false>(tempty_bar(as), 0u);
                if (++as == 2) { as = 0; acc_ph ^= 1u; }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// warp per column: centre row -> fp16 hi / lo pieces of scale * c, |c|^2/2 in rowdot.cu's summation order, info word
__global__ void __launch_bounds__(256) wide_pack_centres_kernel(const float* __restrict__ C_all, int d, const int32_t* __restrict__ col_src,
                                                                const uint32_t* __restrict__ col_info, int Kp, float scale,
                                                                __half* __restrict__ Hi, __half* __restrict__ Lo, float2* __restrict__ tab,
                                                                unsigned int* __restrict__ cmax2h) {
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    const int col = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (col >= Kp) return;
    uint2* oh = reinterpret_cast<uint2*>(Hi + (size_t)col * d);
    uint2* ol = reinterpret_cast<uint2*>(Lo + (size_t)col * d);
    const int src = col_src[col];
    if (src < 0) {
        for (int v = lane; v < nvec; v += 32) { oh[v] = make_uint2(0u, 0u); ol[v] = make_uint2(0u, 0u); }
        if (lane == 0) tab[col] = make_float2(INFINITY, __uint_as_float(col_info[col] | kColPad));
        return;
    }
    const float4* cr = reinterpret_cast<const float4*>(C_all + (size_t)src * d);
    float c2 = 0.f;
    for (int v = lane; v < nvec; v += 32) {
        const float4 a = __ldg(cr + v);
        c2 = fmaf(a.x, a.x, c2); c2 = fmaf(a.y, a.y, c2); c2 = fmaf(a.z, a.z, c2); c2 = fmaf(a.w, a.w, c2);
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
    c2 = warp_sum(c2);
    if (lane == 0) {
        const float h = 0.5f * c2;
        tab[col] = make_float2(h, __uint_as_float(col_info[col]));
        atomicMax(cmax2h, __float_as_uint(h));           // non-negative floats order like their bit patterns
    }
}

struct WideSegs {
    int n;
    int off[kMaxSeg + 1];        // centre slots of labelling L: [off[L], off[L+1])
    int col0[kMaxSeg];           // first column of labelling L in the table of this launch (-1: not in it)
    unsigned int active;
};

// Flagged (row, labelling) pairs -- sign bit of the label word, candidate bits (index & 15) at [8, 24) -- are decided here: exact fp32 scores of the candidate
// centres in rowdot.cu's arithmetic (two FMA chains over lane-strided float4, warp_sum, |c|^2/2 - x.c), first index wins.
// Warps scan 32 labels at a time and serve the flagged ones together.
__global__ void __launch_bounds__(256, 2) wide_refine_kernel(const float* __restrict__ X, int d, const float* __restrict__ C_all,
                                                          const float2* __restrict__ tab, const WideSegs sg,
                                                          int32_t* __restrict__ labels_new, long long label_stride, long long n,
                                                          unsigned int* __restrict__ n_fix) {
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    const bool two_chains = (d % 128 == 0) && (d / 128 <= 8);   // rowdot.cu: register-cached rows alternate two FMA chains
    const long long cpl = (n + 127) / 128;                       // 128-row chunks per labelling: 4 loads in flight per lane
    const long long total = cpl * sg.n;
    unsigned int flagged_total = 0;
    for (long long ch = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); ch < total; ch += (long long)gridDim.x * 8) {
        const int L = (int)(ch / cpl);
        if (sg.col0[L] < 0) continue;                            // not in this launch
        const long long r0 = (ch - (long long)L * cpl) * 128 + lane;
        const long long at0 = (long long)L * label_stride + r0;
        int v4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v4[u] = (r0 + 32 * u < n) ? labels_new[at0 + 32 * u] : 0;
        const int k = sg.off[L + 1] - sg.off[L];
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
            const int v = v4[0];
            v4[0] = v4[1]; v4[1] = v4[2]; v4[2] = v4[3];
            const long long r = r0 + 32 * u, at = at0 + 32 * u;
            unsigned int m = __ballot_sync(0xffffffffu, v < 0);
            if (!m) continue;
            const unsigned int mine = ((unsigned int)v >> 8) & 0xffffu;
            flagged_total += __popc(m);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const unsigned int mask = __shfl_sync(0xffffffffu, mine, src);
                const long long row = r - lane + src;
                const float4* xr = reinterpret_cast<const float4*>(X + (size_t)row * d);
                float4 xc[8];                                    // the row, cached for d <= 1024 (rowdot.cu's DV path)
                if (two_chains) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) xc[i] = (lane + 32 * i < nvec) ? __ldg(xr + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                float best = INFINITY;
                int arg = 0;
                for (int c = 0; c < k; ++c) {
                    if (!((mask >> (c & 15)) & 1u)) continue;
                    const float4* cr = reinterpret_cast<const float4*>(C_all + (size_t)(sg.off[L] + c) * d);
                    float p = 0.f, q2 = 0.f;
                    if (two_chains) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (lane + 32 * i < nvec) {
                                const float4 a = xc[i], b = __ldg(cr + lane + 32 * i);
                                if (i & 1) { q2 = fmaf(a.x, b.x, q2); q2 = fmaf(a.y, b.y, q2); q2 = fmaf(a.z, b.z, q2); q2 = fmaf(a.w, b.w, q2); }
                                else { p = fmaf(a.x, b.x, p); p = fmaf(a.y, b.y, p); p = fmaf(a.z, b.z, p); p = fmaf(a.w, b.w, p); }
                            }
                        }
                    } else {
                        for (int vv = lane; vv < nvec; vv += 32) {
                            const float4 a = __ldg(xr + vv), b = __ldg(cr + vv);
                            p = fmaf(a.x, b.x, p); p = fmaf(a.y, b.y, p); p = fmaf(a.z, b.z, p); p = fmaf(a.w, b.w, p);
                        }
                    }
                    p = warp_sum(p + q2);
                    const float s = tab[sg.col0[L] + c].x - p;
                    if (s < best) { best = s; arg = c; }
                }
                if (lane == src) labels_new[at] = arg;
            }
        }
    }
    if (lane == 0 && flagged_total && n_fix) atomicAdd(n_fix, flagged_total);
}

// grid (x, n_lab): cluster sizes of the new labels and the number of rows whose label changed
__global__ void __launch_bounds__(256) wide_count_kernel(const int32_t* __restrict__ lab_new, long long new_stride,
                                                         const int32_t* __restrict__ lab_old, long long old_stride, long long n,
                                                         const Segs sg, unsigned long long* __restrict__ counts,
                                                         unsigned long long* __restrict__ changed) {
    extern __shared__ unsigned int hist[];
    __shared__ unsigned int s_chg;
    const int L = blockIdx.y;
    if (!((sg.active >> L) & 1u)) return;
    const int k = sg.off[L + 1] - sg.off[L];
    for (int c = threadIdx.x; c < k; c += blockDim.x) hist[c] = 0u;
    if (threadIdx.x == 0) s_chg = 0u;
    __syncthreads();
    const int32_t* a = lab_new + (long long)L * new_stride;
    const int32_t* b = lab_old + (long long)L * old_stride;
    unsigned int chg = 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const int nw = __ldg(a + r);
        chg += (nw != __ldg(b + r)) ? 1u : 0u;
        atomicAdd(&hist[nw], 1u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) chg += __shfl_xor_sync(0xffffffffu, chg, o);
    if ((threadIdx.x & 31) == 0 && chg) atomicAdd(&s_chg, chg);
    __syncthreads();
    for (int c = threadIdx.x; c < k; c += blockDim.x)
        if (hist[c]) atomicAdd(counts + sg.off[L] + c, (unsigned long long)hist[c]);
    if (threadIdx.x == 0 && s_chg) atomicAdd(changed + L, (unsigned long long)s_chg);
}

// grid (x, n_lab): labels_old = labels_new; labellings with at most `cap` changes also list them as {row, L, old, new}
__global__ void __launch_bounds__(256) wide_commit_kernel(const int32_t* __restrict__ lab_new, long long new_stride,
                                                          int32_t* __restrict__ lab_old, long long old_stride, long long n, const Segs sg,
                                                          const unsigned long long* __restrict__ changed, unsigned long long cap,
                                                          int4* __restrict__ items, unsigned int* __restrict__ n_items,
                                                          unsigned int items_cap) {
    const int L = blockIdx.y;
    if (!((sg.active >> L) & 1u)) return;
    const bool list = changed[L] <= cap;
    const int lane = threadIdx.x & 31;
    const int32_t* a = lab_new + (long long)L * new_stride;
    int32_t* b = lab_old + (long long)L * old_stride;
    for (long long r0 = (long long)blockIdx.x * blockDim.x; r0 < n; r0 += (long long)gridDim.x * blockDim.x) {
        const long long r = r0 + threadIdx.x;
        int nw = 0, od = 0;
        bool diff = false;
        if (r < n) { nw = __ldg(a + r); od = b[r]; diff = nw != od; }
        if (diff) b[r] = nw;
        if (list) {
            const unsigned int m = __ballot_sync(0xffffffffu, diff);
            if (m) {
                unsigned int slot = 0;
                if (lane == 0) slot = atomicAdd(n_items, (unsigned int)__popc(m));
                slot = __shfl_sync(0xffffffffu, slot, 0) + (unsigned int)__popc(m & ((1u << lane) - 1u));
                if (diff && slot < items_cap) items[slot] = make_int4((int)r, L, od, nw);
            }
        }
    }
}

__device__ __forceinline__ long long to_fixed(float v, double two_s) { return __double2ll_rn((double)v * two_s); }

// warp per change: acc[new] += x, acc[old] -= x in fixed point (exact; order does not matter)
__global__ void __launch_bounds__(256) wide_delta_kernel(const float* __restrict__ X, int d, const int4* __restrict__ items,
                                                         const unsigned int* __restrict__ n_items, unsigned int items_cap, const Segs sg,
                                                         double two_s, long long* __restrict__ acc) {
    const int lane = threadIdx.x & 31;
    const unsigned int n = min(*n_items, items_cap);
    for (unsigned int it = blockIdx.x * 8 + (threadIdx.x >> 5); it < n; it += gridDim.x * 8) {
        const int4 f = items[it];
        const float* xr = X + (size_t)f.x * d;
        unsigned long long* to = reinterpret_cast<unsigned long long*>(acc + (size_t)(sg.off[f.y] + f.w) * d);
        unsigned long long* from = f.z >= 0 ? reinterpret_cast<unsigned long long*>(acc + (size_t)(sg.off[f.y] + f.z) * d) : nullptr;
        for (int e = lane; e < d; e += 32) {
            const long long qv = to_fixed(__ldg(xr + e), two_s);
            atomicAdd(to + e, (unsigned long long)qv);
            if (from) atomicAdd(from + e, (unsigned long long)(-qv));
        }
    }
}

constexpr int kWSlice = 256;   // dims per CTA slice == threads per CTA

// Sums from scratch for the labellings [L0, L1) that changed too much for the list (changed[L] > cap), fixed point;
// structure of cluster_sums_kernel (accumulate.cu): grid (GX, d/256), thread t owns column t of every cluster slot.
__global__ void __launch_bounds__(kWSlice) wide_full_kernel(const float* __restrict__ X, long long n, int d, const int32_t* __restrict__ labels,
                                                            const Segs sg, int L0, int L1, const unsigned long long* __restrict__ changed,
                                                            unsigned long long cap, double two_s, long long rows_per,
                                                            long long* __restrict__ part) {
    extern __shared__ __align__(16) long long wacc[];   // [kk][256]
    const int t = threadIdx.x;
    unsigned int todo = 0;
    for (int L = L0; L < L1; ++L)
        if (((sg.active >> L) & 1u) && changed[L] > cap) todo |= 1u << L;
    if (!todo) return;
    const int base0 = sg.off[L0], kk = sg.off[L1] - base0;
    const int dim = blockIdx.y * kWSlice + t;
    const bool live = dim < d;
    for (int c = 0; c < kk; ++c) wacc[c * kWSlice + t] = 0ll;
    const long long r0 = (long long)blockIdx.x * rows_per;
    const long long r1 = min(n, r0 + rows_per);
    const float* xp = X + (live ? dim : 0);
    long long r = r0;
    for (; r + 4 <= r1; r += 4) {
        long long v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = live ? to_fixed(__ldg(xp + (r + u) * d), two_s) : 0ll;
        for (int L = L0; L < L1; ++L) {
            if (!((todo >> L) & 1u)) continue;
            const int32_t* lab = labels + L * sg.label_stride + r;
            const int base = sg.off[L] - base0;
#pragma unroll
            for (int u = 0; u < 4; ++u) wacc[(base + __ldg(lab + u)) * kWSlice + t] += v[u];
        }
    }
    for (; r < r1; ++r) {
        const long long v = live ? to_fixed(__ldg(xp + r * d), two_s) : 0ll;
        for (int L = L0; L < L1; ++L) {
            if (!((todo >> L) & 1u)) continue;
            wacc[(sg.off[L] - base0 + __ldg(labels + L * sg.label_stride + r)) * kWSlice + t] += v;
        }
    }
    if (live) {
        long long* out = part + (size_t)blockIdx.x * kk * d;
        for (int c = 0; c < kk; ++c) out[(size_t)c * d + dim] = wacc[c * kWSlice + t];
    }
}

__global__ void __launch_bounds__(256) wide_full_reduce_kernel(const long long* __restrict__ part, int nparts, const Segs sg, int L0, int L1,
                                                               const unsigned long long* __restrict__ changed, unsigned long long cap, int d,
                                                               long long* __restrict__ acc) {
    const int base0 = sg.off[L0], kk = sg.off[L1] - base0;
    const long long len = (long long)kk * d;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    const int c = (int)(e / d) + base0;
    int L = L0;
    while (L + 1 < L1 && c >= sg.off[L + 1]) ++L;
    if (!(((sg.active >> L) & 1u) && changed[L] > cap)) return;
    long long s = 0;
    for (int p = 0; p < nparts; ++p) s += part[(size_t)p * len + e];
    acc[(size_t)base0 * d + e] = s;
}

// kUpdSlices CTAs per labelling (lloyd_update_multi_kernel of accumulate.cu on fixed-point sums): every CTA turns its slice of
// the k x d sums into centres and a partial squared shift; the LAST CTA of a labelling to finish (a counter in global memory)
// adds the partials in slice order -- the same bits whichever CTA that is -- and writes the labelling's stats.  One CTA per
// labelling took 0.11 ms at k = 32, d = 1024 (32 768 int64 -> fp64 divisions on the slow fp64 pipe), 227 times per config-4 fit.
constexpr int kUpdSlices = 8;

__global__ void __launch_bounds__(256) wide_update_kernel(const long long* __restrict__ acc, double inv_two_s, const long long* __restrict__ counts,
                                                          const long long* __restrict__ n_changed, const Segs sg, int d,
                                                          const float* __restrict__ C_old, float* __restrict__ C_new,
                                                          double* __restrict__ stats, int stride, double* __restrict__ partial,
                                                          unsigned int* __restrict__ done) {
    __shared__ double sred[8];
    __shared__ bool s_last;
    const int L = blockIdx.x, slice = blockIdx.y;
    if (!((sg.active >> L) & 1u)) return;
    const int base = sg.off[L], k = sg.off[L + 1] - base;
    int amax = 0;
    long long best = -1;
    for (int c = 0; c < k; ++c) { const long long nn = counts[base + c]; if (nn > best) { best = nn; amax = c; } }
    const int total = k * d;
    const int per = (total + kUpdSlices - 1) / kUpdSlices;
    const int e0 = slice * per, e1 = min(total, e0 + per);
    double shift = 0.0;
    for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int c = e / d, j = e - c * d;
        const long long nn = counts[base + c];
        const int src = nn > 0 ? c : amax;
        const long long ns = nn > 0 ? nn : best;
        const size_t ge = (size_t)(base + c) * d + j;
        const float v = ns > 0 ? (float)(((double)acc[(size_t)(base + src) * d + j] * inv_two_s) / (double)ns) : C_old[ge];
        const double df = (double)v - (double)C_old[ge];
        shift = fma(df, df, shift);
        C_new[ge] = v;
    }
    shift = warp_sum(shift);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = shift;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += sred[w];
        partial[L * kUpdSlices + slice] = s;
        __threadfence();
        s_last = atomicAdd(done + L, 1u) == (unsigned)(kUpdSlices - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double* st = stats + (size_t)L * stride;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int q = 0; q < kUpdSlices; ++q) s += __ldcg(partial + L * kUpdSlices + q);
        st[0] = s;
        st[1] = (double)n_changed[L];
        done[L] = 0u;                                          // ready for the next launch (stream order)
    }
    for (int c = threadIdx.x; c < k; c += blockDim.x) st[2 + c] = (double)counts[base + c];
}

static int wide_gx(int d) {
    const int ds = (d + kWSlice - 1) / kWSlice;
    return max(1, num_sms() * 2 / ds);
}

}  // namespace tc
}  // namespace vtf

using namespace vtf;

static int wide_segs(Segs& sg, int n_lab, const int32_t* k_off_h, uint32_t active, int64_t label_stride, const char* what) {
    if (n_lab <= 0 || n_lab > kMaxSeg || !k_off_h || k_off_h[0] != 0) return fail(VTF_ERR_ARG, what, "1..32 labellings, k_off[0] == 0");
    sg.n = n_lab; sg.active = active; sg.label_stride = label_stride;
    for (int i = 0; i <= n_lab; ++i) sg.off[i] = k_off_h[i];
    return VTF_OK;
}

extern "C" {

int vtf_wide_pack_centres(const float* C_all, int d, const int32_t* col_src, const uint32_t* col_info, int Kp, float scale,
                          void* Bhi, void* Blo, float* tab, uint32_t* cmax2h, void* stream) {
    VTF_REQUIRE(C_all && col_src && col_info && Bhi && Blo && tab && cmax2h && d > 0 && d % 4 == 0 && Kp > 0 && Kp % 256 == 0 &&
                    scale > 0.f,
                "vtf_wide_pack_centres");
    VTF_CUDA(cudaMemsetAsync(cmax2h, 0, sizeof(uint32_t), (cudaStream_t)stream), "vtf_wide_pack_centres");
    VTF_LAUNCH(tc::wide_pack_centres_kernel, (Kp + 7) / 8, 256, 0, stream, C_all, d, col_src, col_info, Kp, scale, (__half*)Bhi,
               (__half*)Blo, (float2*)tab, cmax2h);
    VTF_CHECK_LAST("vtf_wide_pack_centres");
    return VTF_OK;
}

int vtf_wide_assign(const void* Xhi, const void* Xlo, int64_t n_rows, int64_t P, int d, const void* Bhi, const void* Blo, int Kp,
                    int n_cols, const float* tab, const float* rown, const uint32_t* cmax2h, float scale, float eps, int32_t* labels_new,
                    int64_t label_stride, void* stream) {
    VTF_REQUIRE(Xhi && Xlo && Bhi && Blo && tab && rown && cmax2h && labels_new && n_rows >= 0 && P >= n_rows &&
                    P % 256 == 0 && d >= 64 && d <= 4096 && d % 64 == 0 && Kp > 0 && Kp % 256 == 0 && Kp <= tc::kWMaxCols &&
                    n_cols > 0 && n_cols <= Kp && Kp - n_cols < 256 && scale > 0.f && eps >= 0.f && label_stride >= n_rows,
                "vtf_wide_assign (P % 256 == 0, d % 64 == 0, Kp % 256 == 0, Kp <= 2048, Kp - 256 < n_cols <= Kp)");
    if (n_rows == 0) return VTF_OK;
    VTF_REQUIRE(label_stride * 4096 < ((int64_t)1 << 43) && label_stride < ((int64_t)1 << 30) / kMaxSeg,
                "vtf_wide_assign: label_stride * 32 labellings must stay below 2^30");
    int rc = vtf_device_check();
    if (rc != VTF_OK) return rc;
    CUtensorMap mxh, mxl, mch, mcl;
    if ((rc = tc::make_operand_map(&mxh, Xhi, P, d, tc::BM, "vtf_wide_assign")) != VTF_OK) return rc;
    if ((rc = tc::make_operand_map(&mxl, Xlo, P, d, tc::BM, "vtf_wide_assign")) != VTF_OK) return rc;
    if ((rc = tc::make_operand_map(&mch, Bhi, Kp, d, tc::BM, "vtf_wide_assign")) != VTF_OK) return rc;
    if ((rc = tc::make_operand_map(&mcl, Blo, Kp, d, tc::BM, "vtf_wide_assign")) != VTF_OK) return rc;
    const int n_tiles = (n_cols + 255) / 256;
    const int n_last = ((n_cols - (n_tiles - 1) * 256) + 31) / 32 * 32;
    CUtensorMap mchl, mcll;                       // the last tile's centre rows: boxes of n_last / 2 rows
    if ((rc = tc::make_operand_map(&mchl, Bhi, Kp, d, n_last / 2, "vtf_wide_assign")) != VTF_OK) return rc;
    if ((rc = tc::make_operand_map(&mcll, Blo, Kp, d, n_last / 2, "vtf_wide_assign")) != VTF_OK) return rc;
    tc::WideParams prm;
    prm.tab = (const float2*)tab;
    prm.rown = rown;
    prm.cmax2h = cmax2h;
    prm.labels_new = labels_new;
    prm.label_stride = label_stride;
    prm.n_rows = n_rows;
    prm.n_units = (n_rows + 255) / 256;
    prm.n_tiles = n_tiles;
    prm.n_last = n_last;
    prm.n_kb = d / tc::BK;
    prm.inv = 1.f / (scale * scale);
    prm.eps = eps;
    const int grid = 2 * (int)min((int64_t)(num_sms() / 2), (int64_t)prm.n_units);
    const size_t smem = 1024 + (size_t)tc::kWStages * tc::kWStageBytes + 256 + (size_t)Kp * sizeof(float2);
    VTF_CUDA(cudaFuncSetAttribute(tc::wide_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "vtf_wide_assign/smem");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(tc::kWThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VTF_CUDA(cudaLaunchKernelEx(&cfg, tc::wide_assign_kernel, mxh, mxl, mch, mcl, mchl, mcll, prm), "vtf_wide_assign/launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    VTF_CHECK_LAST("vtf_wide_assign");
    return VTF_OK;
}

int vtf_wide_refine(const float* X, int d, const float* C_all, const float* tab, int n_lab, const int32_t* k_off_h,
                    const int32_t* col0_h, int32_t* labels_new, int64_t label_stride, int64_t n, uint32_t* n_fix, void* stream) {
    VTF_REQUIRE(X && C_all && tab && k_off_h && col0_h && labels_new && d > 0 && d % 4 == 0 && n_lab > 0 &&
                    n_lab <= kMaxSeg && n >= 0 && label_stride >= n,
                "vtf_wide_refine");
    if (n == 0) return VTF_OK;
    tc::WideSegs sg = {};
    sg.n = n_lab;
    for (int i = 0; i <= n_lab; ++i) sg.off[i] = k_off_h[i];
    for (int i = 0; i < n_lab; ++i) sg.col0[i] = col0_h[i];
    const int64_t chunks = ((n + 127) / 128) * n_lab;
    const int grid = (int)min((int64_t)num_sms() * 8, (chunks + 7) / 8);
    VTF_LAUNCH(tc::wide_refine_kernel, grid, 256, 0, stream, X, d, C_all, (const float2*)tab, sg, labels_new,
               (long long)label_stride, (long long)n, n_fix);
    VTF_CHECK_LAST("vtf_wide_refine");
    return VTF_OK;
}

int vtf_wide_count(const int32_t* lab_new, int64_t new_stride, const int32_t* lab_old, int64_t old_stride, int64_t n, int n_lab,
                   const int32_t* k_off_h, uint32_t active, int64_t* counts, int64_t* changed, void* stream) {
    VTF_REQUIRE(lab_new && lab_old && counts && changed && n >= 0 && new_stride >= n && old_stride >= n, "vtf_wide_count");
    Segs sg = {};
    int rc = wide_segs(sg, n_lab, k_off_h, active, 0, "vtf_wide_count");
    if (rc != VTF_OK) return rc;
    if (n == 0 || active == 0) return VTF_OK;
    int kmax = 1;
    for (int L = 0; L < n_lab; ++L) kmax = max(kmax, k_off_h[L + 1] - k_off_h[L]);
    VTF_REQUIRE((size_t)kmax * 4 <= 48 * 1024, "vtf_wide_count: k too large");
    const int gx = (int)max((int64_t)1, min((int64_t)(num_sms() * 8 / max(1, __builtin_popcount(active)) + 1), (n + 255) / 256));
    VTF_LAUNCH(tc::wide_count_kernel, dim3(gx, n_lab), 256, (size_t)kmax * 4, stream, lab_new, (long long)new_stride, lab_old,
               (long long)old_stride, (long long)n, sg, (unsigned long long*)counts, (unsigned long long*)changed);
    VTF_CHECK_LAST("vtf_wide_count");
    return VTF_OK;
}

int vtf_wide_commit(const int32_t* lab_new, int64_t new_stride, int32_t* lab_old, int64_t old_stride, int64_t n, int n_lab,
                    const int32_t* k_off_h, uint32_t active, const int64_t* changed, int64_t cap, void* items, uint32_t* n_items,
                    uint32_t items_cap, void* stream) {
    VTF_REQUIRE(lab_new && lab_old && changed && items && n_items && n >= 0 && new_stride >= n && old_stride >= n && cap >= 0,
                "vtf_wide_commit");
    Segs sg = {};
    int rc = wide_segs(sg, n_lab, k_off_h, active, 0, "vtf_wide_commit");
    if (rc != VTF_OK) return rc;
    VTF_CUDA(cudaMemsetAsync(n_items, 0, sizeof(uint32_t), (cudaStream_t)stream), "vtf_wide_commit");
    if (n == 0 || active == 0) return VTF_OK;
    const int gx = (int)max((int64_t)1, min((int64_t)(num_sms() * 8 / max(1, __builtin_popcount(active)) + 1), (n + 255) / 256));
    VTF_LAUNCH(tc::wide_commit_kernel, dim3(gx, n_lab), 256, 0, stream, lab_new, (long long)new_stride, lab_old, (long long)old_stride,
               (long long)n, sg, (const unsigned long long*)changed, (unsigned long long)cap, (int4*)items, n_items, items_cap);
    VTF_CHECK_LAST("vtf_wide_commit");
    return VTF_OK;
}

int vtf_wide_sums_delta(const float* X, int d, const void* items, const uint32_t* n_items, uint32_t items_cap, int n_lab,
                        const int32_t* k_off_h, double two_s, int64_t* acc, void* stream) {
    VTF_REQUIRE(X && items && n_items && acc && d > 0 && two_s > 0.0, "vtf_wide_sums_delta");
    Segs sg = {};
    int rc = wide_segs(sg, n_lab, k_off_h, 0xffffffffu, 0, "vtf_wide_sums_delta");
    if (rc != VTF_OK) return rc;
    if (items_cap == 0) return VTF_OK;
    const int grid = (int)min((int64_t)num_sms() * 8, (int64_t)(items_cap + 7) / 8);
    VTF_LAUNCH(tc::wide_delta_kernel, grid, 256, 0, stream, X, d, (const int4*)items, n_items, items_cap, sg, two_s, (long long*)acc);
    VTF_CHECK_LAST("vtf_wide_sums_delta");
    return VTF_OK;
}

size_t vtf_wide_sums_ws_bytes(int k_group, int d) {
    return (size_t)tc::wide_gx(d) * (size_t)k_group * (size_t)d * sizeof(long long);
}

int vtf_wide_sums_full(const float* X, int64_t n, int d, const int32_t* labels, int64_t label_stride, int n_lab,
                       const int32_t* k_off_h, uint32_t active, int L0, int L1, const int64_t* changed, int64_t cap, double two_s,
                       int64_t* acc, void* ws, void* stream) {
    VTF_REQUIRE(X && labels && changed && acc && ws && n >= 0 && d > 0 && d % 4 == 0 && label_stride >= n && L0 >= 0 && L1 > L0 &&
                    L1 <= n_lab && two_s > 0.0 && cap >= -1,
                "vtf_wide_sums_full");
    Segs sg = {};
    int rc = wide_segs(sg, n_lab, k_off_h, active, label_stride, "vtf_wide_sums_full");
    if (rc != VTF_OK) return rc;
    const int kk = k_off_h[L1] - k_off_h[L0];
    const size_t smem = (size_t)kk * tc::kWSlice * sizeof(long long);
    VTF_REQUIRE(smem <= 200 * 1024, "vtf_wide_sums_full: more than 100 cluster slots in one group");
    if (n == 0) return VTF_OK;
    const int gx = tc::wide_gx(d);
    const int ds = (d + tc::kWSlice - 1) / tc::kWSlice;
    const long long rows_per = (n + gx - 1) / gx;
    // cap = -1: every active labelling of the group is recomputed
    const unsigned long long ucap = cap < 0 ? 0ull : (unsigned long long)cap;
    VTF_CUDA(cudaFuncSetAttribute(tc::wide_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "vtf_wide_sums_full");
    VTF_LAUNCH(tc::wide_full_kernel, dim3(gx, ds), tc::kWSlice, smem, stream, X, (long long)n, d, labels, sg, L0, L1,
               (const unsigned long long*)changed, ucap, two_s, rows_per, (long long*)ws);
    VTF_CHECK_LAST("vtf_wide_sums_full");
    const long long len = (long long)kk * d;
    VTF_LAUNCH(tc::wide_full_reduce_kernel, (unsigned)((len + 255) / 256), 256, 0, stream, (const long long*)ws, gx, sg, L0, L1,
               (const unsigned long long*)changed, ucap, d, (long long*)acc);
    VTF_CHECK_LAST("vtf_wide_sums_full/reduce");
    return VTF_OK;
}

int vtf_wide_update(const int64_t* acc, double inv_two_s, const int64_t* counts, const int64_t* changed, int n_lab,
                    const int32_t* k_off_h, uint32_t active, int d, const float* C_old, float* C_new, double* stats,
                    int stats_stride, void* stream) {
    VTF_REQUIRE(acc && counts && changed && C_old && C_new && stats && d > 0 && stats_stride >= 2 && inv_two_s > 0.0, "vtf_wide_update");
    Segs sg = {};
    int rc = wide_segs(sg, n_lab, k_off_h, active, 0, "vtf_wide_update");
    if (rc != VTF_OK) return rc;
    if (active == 0) return VTF_OK;
    // scratch of the slice reduction: kMaxSeg * kUpdSlices partial shifts + kMaxSeg counters (zero between launches), allocated once;
    // launches are stream-ordered (one fit at a time per process)
    static double* scratch = nullptr;
    if (!scratch) {
        const size_t bytes = (size_t)kMaxSeg * tc::kUpdSlices * sizeof(double) + (size_t)kMaxSeg * sizeof(unsigned int);
        VTF_CUDA(cudaMalloc(&scratch, bytes), "vtf_wide_update/scratch");
        VTF_CUDA(cudaMemset(scratch, 0, bytes), "vtf_wide_update/scratch");
    }
    unsigned int* done = reinterpret_cast<unsigned int*>(scratch + (size_t)kMaxSeg * tc::kUpdSlices);
    VTF_LAUNCH(tc::wide_update_kernel, dim3(n_lab, tc::kUpdSlices), 256, 0, stream, (const long long*)acc, inv_two_s, (const long long*)counts,
               (const long long*)changed, sg, d, C_old, C_new, stats, stats_stride, scratch, done);
    VTF_CHECK_LAST("vtf_wide_update");
    return VTF_OK;
}

}  // extern "C"
