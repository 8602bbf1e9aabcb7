// Rows against a small matrix, register-tiled: the fast form of the Lloyd E-step (all labellings of a k-range in one
// pass, SK/cluster/_k_means_lloyd.pyx:168-218) and of nearest-reference classification (grouping.py:51-56) for up
// to kMaxC = 32 centres per pass.
//
// One thread owns TWO rows and keeps 2 x (up to 32) fp32 dot-product accumulators in registers, so there is no
// cross-lane reduction at all (the warp-per-row kernel in rowdot.cu spends more issue slots on shuffles than on FMAs
// once sum(k) reaches 20).  X streams through shared memory in [512 rows x 32 floats] stages filled by TMA (two
// boxes of 256 rows x 128 bytes per stage, SWIZZLE_128B, so that lane = row reads are bank-conflict free; rows past
// the end arrive as zeros); the centres sit in shared memory and are read as warp-wide broadcasts.
// What the copy granularity costs (1M x 512, k = 6 / sum(k) = 20 / classify R = 32, ms): 64-byte row pieces starve
// both copy paths -- cp.async 0.63 / 1.08 / 1.21 with 2..5 stages and any L2 prefetch hint, TMA 0.53 / 0.97 / 1.14 --
// because the request rate, not the bytes in flight, is the limit; 128-byte pieces: cp.async 0.44 / 0.85 / 1.07,
// TMA 0.39 / 0.86 / 0.98 (this version; rowdot.cu: 0.69 / 2.09 / 2.34).  k = 6 then runs at 5.2 TB/s; from sum(k) ~ 20
// the shared-memory pipe is the limit (a warp-wide broadcast LDS.128 still takes a cycle per quarter-warp).
// Bound: one HBM read of X next to fp32 FMA issue (2 N sum(k) d flop); measured in tools/kernel_bench.py.
#include "tc_common.cuh"

namespace vtf {

constexpr int kMaxC = 32;          // centres per pass = accumulators per row
constexpr int kRows = 2;           // rows per thread (4 rows x 128 threads: 1.34 ms, one warp per scheduler cannot hide the latencies)
constexpr int kChunk = 32;         // floats of every row per stage (128 bytes: smaller pieces starve the copy engines)
constexpr int kTileThreads = 256;  // 512-row tiles, 64 KB per stage
constexpr int kMaxStages = 3;      // as many 64 KB stages as fit next to the centres (2 for sum(k) = 20 at d = 512)

struct TileList {
    int n_act;                     // active centres, compacted
    int n_seg;                     // active labellings
    short cidx[kMaxC];             // compacted position -> centre row in C_all
    short seg_end[kMaxSeg];        // compacted position one past the last centre of active labelling s
    short seg_id[kMaxSeg];         // its labelling index (label plane, n_changed slot)
    short seg_base[kMaxSeg];       // first centre row of that labelling in C_all (label = centre row - base)
};

struct TileEpi {
    int32_t* labels;               // Lloyd: [n_lab][label_stride]
    long long label_stride;
    unsigned long long* counts;    // Lloyd: [Ktot]
    unsigned long long* n_changed; // Lloyd: [n_lab]
    float* dist;                   // classify: [N, k]
    int64_t* label;                // classify: [N]
    float thr;                     // classify: < 0 or NaN = disabled
};

// MODE 0: Lloyd (score = |c|^2 - 2 x.c per labelling), MODE 1: classify (score = clip(1 - x.c, 0, 2))
// NT threads, each owning kRows rows of a tile of kRows * NT rows (thread t: rows t, t + NT, ...).  NG = groups of 4
// centres (the centre count rounded up to a multiple of 4; padding centres are zero rows whose scores are never used).
// n_stages (2..5) stages of X, n_stages - 1 of them in flight.
template <int MODE, int NG>
__global__ void __launch_bounds__(kTileThreads, 1) rowtile_kernel(const __grid_constant__ CUtensorMap map_x, int64_t N, int d,
                                                                  const float* __restrict__ C, int k_total, int n_stages,
                                                                  const TileList tl, const TileEpi ep) {
    using namespace tc;
    constexpr int NT = kTileThreads;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    // the 64-byte swizzle works on absolute shared-memory address bits: stages start on a 1024-byte boundary
    unsigned char* smem_raw = smem_dyn + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_dyn) & 1023u)) & 1023u);
    constexpr int TR = kRows * NT;
    constexpr int kRowBytes = kChunk * 4;                                              // 128
    constexpr int kStageBytes = TR * kRowBytes;
    constexpr int NC = 4 * NG;                                                         // accumulators per row
    float* sC = reinterpret_cast<float*>(smem_raw + n_stages * kStageBytes);           // [NC][d]
    float* sC2 = sC + (size_t)NC * d;                                                  // [kMaxC]
    unsigned int* s_counts = reinterpret_cast<unsigned int*>(sC2 + kMaxC);             // [kMaxC] by compacted position
    unsigned int* s_changed = s_counts + kMaxC;                                        // [kMaxSeg] by active labelling
    const uint32_t stage_u = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(s_changed + kMaxSeg);     // [kMaxStages] "stage full" mbarriers
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nvec = d >> 2;
    const int n_chunks = d / kChunk;

    for (int j = warp; j < tl.n_act; j += NT / 32) {                                   // centres -> shared memory
        const float4* src = reinterpret_cast<const float4*>(C + (size_t)tl.cidx[j] * d);
        float4* dst = reinterpret_cast<float4*>(sC + (size_t)j * d);
        float s = 0.f;
        for (int v = lane; v < nvec; v += 32) {
            const float4 a = __ldg(src + v);
            dst[v] = a;
            s = fmaf(a.x, a.x, s); s = fmaf(a.y, a.y, s); s = fmaf(a.z, a.z, s); s = fmaf(a.w, a.w, s);
        }
        // |c|^2 in fp32 (SK/cluster/_k_means_lloyd.pyx:99), same lane-strided order as rowdot.cu
        s = warp_sum(s);
        if (lane == 0) sC2[j] = s;
    }
    for (int i = tid; i < (NC - tl.n_act) * nvec; i += NT)                             // padding centres: zero rows
        reinterpret_cast<float4*>(sC + (size_t)tl.n_act * d)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < kMaxC) s_counts[tid] = 0;
    if (tid < kMaxSeg) s_changed[tid] = 0;
    if (tid == 0) {
        for (int s = 0; s < kMaxStages; ++s) mbar_init(bars + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int64_t n_tiles = (N + TR - 1) / TR;
    // flattened (tile, chunk) sequence of this CTA; step i uses stage i % n_stages
    const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int64_t steps = my_tiles * n_chunks;
    // 128-byte rows, 8 granules of 16 bytes; SWIZZLE_128B puts granule g of row r at position g ^ (r & 7): the eight
    // rows a quarter-warp reads with one LDS.128 then fall into eight different 16-byte bank groups
    auto issue = [&](int64_t step) {                                                   // thread 0 only
        if (step >= steps) return;
        const int64_t tile = blockIdx.x + (step / n_chunks) * gridDim.x;
        const int chunk = (int)(step % n_chunks);
        const uint32_t slot = (uint32_t)(step % n_stages);
        const uint32_t dst = stage_u + slot * kStageBytes, bar = bars + 8u * slot;
        mbar_expect_tx(bar, (uint32_t)kStageBytes);
#pragma unroll
        for (int b = 0; b < TR / 256; ++b)                                             // boxes of 256 rows x 32 floats
            tma_load_2d(&map_x, bar, dst + (uint32_t)b * 256u * kRowBytes, chunk * kChunk, (int)(tile * TR + b * 256));
    };
    if (tid == 0)
        for (int s = 0; s < n_stages; ++s) issue(s);

    float acc[kRows][NC];
    for (int64_t step = 0; step < steps; ++step) {
        const int chunk = (int)(step % n_chunks);
        if (chunk == 0) {
#pragma unroll
            for (int r = 0; r < kRows; ++r)
#pragma unroll
                for (int j = 0; j < NC; ++j) acc[r][j] = 0.f;
        }
        if (step > 0) {
            __syncthreads();                                                           // step - 1 fully consumed: refill its stage
            if (tid == 0) issue(step - 1 + n_stages);
        }
        mbar_wait(bars + 8u * (uint32_t)(step % n_stages), (uint32_t)((step / n_stages) & 1));
        const unsigned char* st = smem_raw + (size_t)(step % n_stages) * kStageBytes;
        const float* cbase = sC + chunk * kChunk;
#pragma unroll
        for (int v = 0; v < kChunk / 4; ++v) {
            float4 x[kRows];
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const int row = tid + NT * r;
                x[r] = *reinterpret_cast<const float4*>(st + row * kRowBytes + ((v ^ (row & 7)) << 4));
            }
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const float4 c = *reinterpret_cast<const float4*>(cbase + (size_t)j * d + 4 * v);         // warp-wide broadcast
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    acc[r][j] = fmaf(x[r].x, c.x, acc[r][j]); acc[r][j] = fmaf(x[r].y, c.y, acc[r][j]);
                    acc[r][j] = fmaf(x[r].z, c.z, acc[r][j]); acc[r][j] = fmaf(x[r].w, c.w, acc[r][j]);
                }
            }
        }
        if (chunk != n_chunks - 1) continue;

        // ---- all of d seen for this tile: scores -> labels
        const int64_t tile = blockIdx.x + (step / n_chunks) * gridDim.x;
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int64_t row = tile * TR + tid + NT * r;
            const bool ok = row < N;
            if (MODE == 0) {
                int s = 0;
                float best = INFINITY;
                int arg = 0;
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    if (j < tl.n_act) {
                        const float sc = fmaf(-2.f, acc[r][j], sC2[j]);
                        if (sc < best) { best = sc; arg = j; }                         // strict '<': first index wins
                        if (j + 1 == tl.seg_end[s]) {
                            // last centre of active labelling s: publish the label, count a change, restart the argmin
                            if (ok) {
                                int32_t* plane = ep.labels + (size_t)tl.seg_id[s] * ep.label_stride;
                                const int lab = tl.cidx[arg] - tl.seg_base[s];
                                if (plane[row] != lab) atomicAdd(&s_changed[s], 1u);
                                plane[row] = lab;
                                atomicAdd(&s_counts[arg], 1u);
                            }
                            best = INFINITY;
                            ++s;
                        }
                    }
                }
            } else {
                float best = INFINITY;
                int arg = 0;
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    acc[r][j] = fminf(fmaxf(1.f - acc[r][j], 0.f), 2.f);
                    if (j < tl.n_act && acc[r][j] < best) { best = acc[r][j]; arg = j; }
                }
                if (ok) {
                    // the distance log [N, k]: 16-byte stores where the row pitch allows it
                    const bool vec = (k_total & 3) == 0;
#pragma unroll
                    for (int g4 = 0; g4 < NG; ++g4) {
                        if (vec && 4 * g4 + 3 < k_total) {
                            *reinterpret_cast<float4*>(ep.dist + row * k_total + 4 * g4) =
                                make_float4(acc[r][4 * g4], acc[r][4 * g4 + 1], acc[r][4 * g4 + 2], acc[r][4 * g4 + 3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (4 * g4 + e < k_total) ep.dist[row * k_total + 4 * g4 + e] = acc[r][4 * g4 + e];
                        }
                    }
                    const bool on = ep.thr >= 0.f;                                     // false for NaN as well
                    ep.label[row] = (on && best >= ep.thr) ? (int64_t)k_total : (int64_t)arg;
                }
            }
        }
    }
    if (MODE == 0) {
        __syncthreads();
        if (tid < tl.n_act && s_counts[tid]) atomicAdd(ep.counts + tl.cidx[tid], (unsigned long long)s_counts[tid]);
        if (tid < tl.n_seg && s_changed[tid]) atomicAdd(ep.n_changed + tl.seg_id[tid], (unsigned long long)s_changed[tid]);
    }
}

static size_t rowtile_smem(int n_act, int d, int stages) {
    const int nc = (n_act + 3) & ~3;
    return (size_t)stages * kRows * kTileThreads * kChunk * 4 + (size_t)nc * d * 4 + (kMaxC + kMaxC + kMaxSeg) * 4 + kMaxStages * 8 + 1024;
}

// X [N, d] fp32 row-major, boxes of 32 floats x 256 rows, 128-byte swizzle, out-of-range rows read as zeros
static int make_rows_map(CUtensorMap* map, const float* X, int64_t N, int d, const char* who) {
    tc::EncodeTiledFn enc = tc::get_encode();
    if (!enc) return fail(VTF_ERR_CUDA, who, "cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)d * 4};
    cuuint32_t box[2] = {(cuuint32_t)kChunk, 256u};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[64];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d)", (int)r);
        return fail(VTF_ERR_CUDA, who, buf);
    }
    return VTF_OK;
}

template <int MODE, int NG>
static int launch_rowtile_ng(const CUtensorMap& mx, int64_t N, int d, const float* C, int k_total, int stages, const TileList& tl,
                             const TileEpi& ep, int grid, size_t smem, void* stream, const char* what) {
    VTF_CUDA(cudaFuncSetAttribute(rowtile_kernel<MODE, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), what);
    VTF_LAUNCH((rowtile_kernel<MODE, NG>), grid, kTileThreads, smem, stream, mx, N, d, C, k_total, stages, tl, ep);
    VTF_CHECK_LAST(what);
    return 0;
}

// 0 = launched, 1 = shape not covered (caller falls back to rowdot.cu), < 0 = error
template <int MODE>
static int launch_rowtile(const float* X, int64_t N, int d, const float* C, int k_total, const TileList& tl, const TileEpi& ep,
                          void* stream, const char* what) {
    if (tl.n_act <= 0 || tl.n_act > kMaxC || d % kChunk != 0 || k_total > 32767 || (reinterpret_cast<uintptr_t>(X) & 15)) return 1;
    if (vtf_device_check() != VTF_OK) return 1;
    const size_t cap = 220 * 1024;
    int stages = kMaxStages;
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("VTF_ROWTILE_STAGES"); forced = e ? atoi(e) : 0; }        // A/B runs
    if (forced >= 2 && forced <= kMaxStages) stages = forced;
    while (stages > 2 && rowtile_smem(tl.n_act, d, stages) > cap) --stages;
    const size_t smem = rowtile_smem(tl.n_act, d, stages);
    if (smem > cap) return 1;
    const int64_t n_tiles = (N + kRows * kTileThreads - 1) / (kRows * kTileThreads);
    const int grid = (int)min((int64_t)num_sms(), n_tiles);
    const int ng = (tl.n_act + 3) / 4;
    CUtensorMap mx;
    const int mrc = make_rows_map(&mx, X, N, d, what);
    if (mrc != VTF_OK) return mrc;
#define VTF_RT_CASE(NGv) \
    case NGv: return launch_rowtile_ng<MODE, NGv>(mx, N, d, C, k_total, stages, tl, ep, grid, smem, stream, what);
    switch (ng) {
        VTF_RT_CASE(1) VTF_RT_CASE(2) VTF_RT_CASE(3) VTF_RT_CASE(4) VTF_RT_CASE(5) VTF_RT_CASE(6) VTF_RT_CASE(7) VTF_RT_CASE(8)
    }
#undef VTF_RT_CASE
    return 1;
}

// ---- entry points used by rowdot.cu's extern "C" functions
int rowtile_lloyd(const float* X, int64_t N, int d, const float* C_all, const Segs& sg, int32_t* labels,
                  unsigned long long* counts, unsigned long long* n_changed, void* stream, const char* what) {
    TileList tl = {};
    for (int L = 0; L < sg.n; ++L) {
        if (!((sg.active >> L) & 1u)) continue;
        for (int c = sg.off[L]; c < sg.off[L + 1]; ++c) {
            if (tl.n_act >= kMaxC) return 1;
            tl.cidx[tl.n_act++] = (short)c;
        }
        tl.seg_end[tl.n_seg] = (short)tl.n_act;
        tl.seg_id[tl.n_seg] = (short)L;
        tl.seg_base[tl.n_seg] = (short)sg.off[L];
        ++tl.n_seg;
    }
    TileEpi ep = {};
    ep.labels = labels; ep.label_stride = sg.label_stride; ep.counts = counts; ep.n_changed = n_changed;
    return launch_rowtile<0>(X, N, d, C_all, sg.off[sg.n], tl, ep, stream, what);
}

int rowtile_classify(const float* X, int64_t N, int d, const float* R, int nR, float thr, float* dist, int64_t* label,
                     void* stream, const char* what) {
    if (nR > kMaxC) return 1;
    TileList tl = {};
    tl.n_act = nR; tl.n_seg = 1;
    for (int c = 0; c < nR; ++c) tl.cidx[c] = (short)c;
    tl.seg_end[0] = (short)nR;
    TileEpi ep = {};
    ep.dist = dist; ep.label = label; ep.thr = thr;
    return launch_rowtile<1>(X, N, d, R, nR, tl, ep, stream, what);
}

}  // namespace vtf
