// Segmented reductions over the rows of X, organised by CELLS.
//
// grouping.py:99-101 fits KMeans once per k of the range; all fits read the same rows.  A row's labels over the L
// fits of a range form a tuple; rows that share the tuple form a *cell* (k = 2..6: at most 720 tuples, 8-20 of them
// occupied on the bench data).  Inside a cell every labelling sends the row to one fixed cluster, so the M-step
// accumulation of SK/cluster/_k_means_lloyd.pyx:215-218
//
//     for every labelling L:  sums[off[L] + label_L(row)] += x_row          (5 read-modify-writes per element)
//
// becomes      acc_cell += x_row  in REGISTERS   (one add per element),  then once per cell  sums[...] += acc_cell.
//
// cell_sums_kernel: persistent CTAs claim chunks of 1024 consecutive rows from a counter; a chunk is grouped by cell with a shared-memory counting sort (keys,
// histogram, scan, scatter of the row numbers), then walks the cells: thread t owns two columns, streams the rows of the
// cell with 16 independent 8-byte loads in flight, adds them up in 64-bit FIXED POINT (round(x * 2^S), the unit of
// lloyd_tc.cu) and folds the cell's total into the per-cluster accumulators in shared memory.  Integer adds are exact
// and order-independent: the result does not depend on the order of the scatter, of the CTAs or of the ranks, and it
// is bit-identical to vtf_wide_sums_full / vtf_wide_sums_delta.  HBM-bound: N*d*4 bytes, read once, 2 KB per row.
//
// The kernel this replaces (cluster_sums_kernel, accumulate.cu) makes one fp64 shared-memory read-modify-write per element
// and labelling: 0.14 of the HBM roofline for k = 2..6 (profiles/r01_kernel_bench_1m.json).
#include "common.cuh"

namespace vtf {

constexpr int kCellKeys = 1024;      // joint label tuples one launch can tell apart
constexpr int kCellRows = 1024;      // rows per chunk; persistent CTAs claim chunks from a counter
constexpr int kCellThreads = 256;
constexpr int kCellW = 2 * kCellThreads;   // columns per CTA slice: thread t owns columns 2t, 2t+1
constexpr int kCellMaxLab = 8;
constexpr int kCellMaxSlots = 48;    // cluster slots of one launch (48 * 512 * 8 B = 192 KB of shared accumulators)
constexpr int kCellUnroll = 16;

struct CellPlan {
    int n_lab;                    // labellings that make up the key (the active ones of the group)
    int row_of[kCellMaxLab];      // labelling index L (row of the label matrix)
    int slot0[kCellMaxLab];       // first cluster slot of L, relative to the group's first slot
    int kk[kCellMaxLab];          // clusters of L
    int mult[kCellMaxLab];        // key = sum_q label_q * mult[q]
    int n_slots;                  // cluster slots of the group (accumulator rows in shared memory)
    long long label_stride;
};

// streaming 8-byte load through the intrinsic (not inline asm): the scheduler knows it is a long-latency load and issues
// the whole unrolled batch before the first use, PROVIDED the launch bounds tell ptxas that only 2 CTAs fit an SM anyway
// (shared memory): without the hint it squeezed the kernel into 32 registers and issued the loads two at a time
__device__ __forceinline__ float2 ldg_stream2(const float2* p) { return __ldcs(p); }

__global__ void __launch_bounds__(kCellThreads, 2) cell_sums_kernel(const float* __restrict__ X, long long n, int d,
                                                                 const int32_t* __restrict__ labels, const CellPlan pl,
                                                                 float two_s, unsigned long long* __restrict__ acc,
                                                                 unsigned int* __restrict__ bad, unsigned int* __restrict__ claim) {
    extern __shared__ __align__(16) unsigned char cell_smem[];
    long long* sacc = reinterpret_cast<long long*>(cell_smem);                               // [n_slots][512]
    unsigned int* cursor = reinterpret_cast<unsigned int*>(sacc + (size_t)pl.n_slots * kCellW);   // [1024] counts, then write cursors
    unsigned int* start = cursor + kCellKeys;                                                // [1024] first position of the cell
    unsigned short* keys = reinterpret_cast<unsigned short*>(start + kCellKeys);             // [2048]
    unsigned short* order = keys + kCellRows;                                                // [2048] rows grouped by cell
    unsigned short* segkey = order + kCellRows;                                              // [1024] occupied cells
    __shared__ unsigned int warp_tot[kCellThreads / 32];
    __shared__ int s_nseg;
    __shared__ unsigned int s_chunk;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int dim = blockIdx.y * kCellW + 2 * t;
    const bool live = dim < d;
    const long long n_chunks = (n + kCellRows - 1) / kCellRows;

    for (int i = t; i < pl.n_slots * kCellW; i += kCellThreads) sacc[i] = 0ll;
    unsigned int nbad = 0;
    for (;;) {
    if (t == 0) s_chunk = atomicAdd(claim + blockIdx.y, 1u);     // chunks in any order: the sums are integers
    for (int i = t; i < kCellKeys; i += kCellThreads) cursor[i] = 0u;
    __syncthreads();
    if ((long long)s_chunk >= n_chunks) break;
    const long long row0 = (long long)s_chunk * kCellRows;
    const int R = (int)min((long long)kCellRows, n - row0);

    // ---- keys + histogram
    for (int i = t; i < R; i += kCellThreads) {
        int key = 0;
        for (int q = 0; q < pl.n_lab; ++q) {
            int l = __ldg(labels + (long long)pl.row_of[q] * pl.label_stride + row0 + i);
            if ((unsigned)l >= (unsigned)pl.kk[q]) { ++nbad; l = 0; }
            key += l * pl.mult[q];
        }
        keys[i] = (unsigned short)key;
        atomicAdd(&cursor[key], 1u);
    }
    __syncthreads();

    // ---- exclusive scan over the 1024 cells (4 per thread): positions in the low half, occupied-cell rank in the high half
    unsigned int c4[4];
    unsigned int local = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { c4[j] = cursor[4 * t + j]; local += c4[j] + (c4[j] ? 0x10000u : 0u); }
    unsigned int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned int base = incl - local;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    if (t == kCellThreads - 1) s_nseg = (int)((base + local) >> 16);
    {
        unsigned int pos = base & 0xffffu, seg = base >> 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            start[4 * t + j] = pos;
            cursor[4 * t + j] = pos;
            if (c4[j]) segkey[seg++] = (unsigned short)(4 * t + j);
            pos += c4[j];
        }
    }
    __syncthreads();

    // ---- scatter the row numbers (any order inside a cell: the sums are integers)
    for (int i = t; i < R; i += kCellThreads) order[atomicAdd(&cursor[keys[i]], 1u)] = (unsigned short)i;
    __syncthreads();

    // ---- walk the cells
    const int nseg = s_nseg;
    const float* xcol = X + (size_t)row0 * d + (live ? dim : 0);
    for (int s = 0; s < nseg; ++s) {
        const int key = segkey[s];
        const int st = (int)start[key], en = (int)cursor[key];
        long long a0 = 0, a1 = 0;
        int j = st;
        for (; j + kCellUnroll <= en; j += kCellUnroll) {          // whole blocks: 16 independent loads, no guards
            float2 v[kCellUnroll];
#pragma unroll
            for (int u = 0; u < kCellUnroll; ++u)
                v[u] = ldg_stream2(reinterpret_cast<const float2*>(xcol + (size_t)order[j + u] * d));
#pragma unroll
            for (int u = 0; u < kCellUnroll; ++u) {
                a0 += __float2ll_rn(v[u].x * two_s);                // power-of-two scaling: exact, then round to nearest even
                a1 += __float2ll_rn(v[u].y * two_s);
            }
        }
        for (; j + 4 <= en; j += 4) {                               // tail of the cell in blocks of 4 ...
            float2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ldg_stream2(reinterpret_cast<const float2*>(xcol + (size_t)order[j + u] * d));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a0 += __float2ll_rn(v[u].x * two_s);
                a1 += __float2ll_rn(v[u].y * two_s);
            }
        }
        for (; j < en; ++j) {                                       // ... and single rows
            const float2 v = ldg_stream2(reinterpret_cast<const float2*>(xcol + (size_t)order[j] * d));
            a0 += __float2ll_rn(v.x * two_s);
            a1 += __float2ll_rn(v.y * two_s);
        }
        if (!live) { a0 = 0; a1 = 0; }
        for (int q = 0; q < pl.n_lab; ++q) {
            const int l = (key / pl.mult[q]) % pl.kk[q];
            long long* dst = sacc + (size_t)(pl.slot0[q] + l) * kCellW + 2 * t;     // this thread's own columns: no conflict
            dst[0] += a0;
            dst[1] += a1;
        }
    }
    __syncthreads();                                                // s_chunk, cursor, order are rewritten by the next claim
    }
    if (nbad && bad) atomicAdd(bad, nbad);

    // ---- fold into the global accumulators (exact integer atomics: any order)
    if (live) {
        for (int c = 0; c < pl.n_slots; ++c) {
            const long long v0 = sacc[(size_t)c * kCellW + 2 * t], v1 = sacc[(size_t)c * kCellW + 2 * t + 1];
            if (v0) atomicAdd(acc + (size_t)c * d + dim, (unsigned long long)v0);
            if (v1) atomicAdd(acc + (size_t)c * d + dim + 1, (unsigned long long)v1);
        }
    }
}

}  // namespace vtf

using namespace vtf;

extern "C" {

int vtf_cell_sums(const float* X, int64_t n, int d, const int32_t* labels, int64_t label_stride, int n_lab,
                  const int32_t* k_off_h, uint32_t active, double two_s, int64_t* acc, uint32_t* bad, void* ws, void* stream) {
    VTF_REQUIRE(X && labels && k_off_h && acc && ws && n >= 0 && d > 0 && d % 2 == 0 && n_lab > 0 && n_lab <= kMaxSeg &&
                    label_stride >= n && k_off_h[0] == 0 && two_s > 0.0 && two_s < 3.0e38,
                "vtf_cell_sums");
    cudaStream_t st = (cudaStream_t)stream;
    // the active labellings start from zero; frozen ones keep their sums
    for (int L = 0; L < n_lab; ++L) {
        if (!((active >> L) & 1u)) continue;
        const int k = k_off_h[L + 1] - k_off_h[L];
        VTF_REQUIRE(k > 0 && k <= kCellKeys && k <= kCellMaxSlots, "vtf_cell_sums: 1 <= k <= 48 per labelling");
        VTF_CUDA(cudaMemsetAsync(acc + (size_t)k_off_h[L] * d, 0, (size_t)k * d * sizeof(int64_t), st), "vtf_cell_sums/zero");
    }
    if (bad) VTF_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t), st), "vtf_cell_sums/zero");
    if (n == 0 || active == 0) return VTF_OK;
    // groups of active labellings: at most 8 of them, at most 1024 label tuples, at most 48 cluster slots from the first to
    // the last labelling of the group (frozen labellings in between keep their slots but take no part in the key)
    int L = 0;
    while (L < n_lab) {
        if (!((active >> L) & 1u)) { ++L; continue; }
        CellPlan pl = {};
        const int L0 = L;
        long long keys = 1;
        int last_end = k_off_h[L0];
        while (L < n_lab) {
            if (!((active >> L) & 1u)) { ++L; continue; }
            const int k = k_off_h[L + 1] - k_off_h[L];
            if (pl.n_lab > 0 && (pl.n_lab >= kCellMaxLab || keys * k > kCellKeys || k_off_h[L + 1] - k_off_h[L0] > kCellMaxSlots)) break;
            pl.row_of[pl.n_lab] = L;
            pl.slot0[pl.n_lab] = k_off_h[L] - k_off_h[L0];
            pl.kk[pl.n_lab] = k;
            pl.mult[pl.n_lab] = (int)keys;
            keys *= k;
            last_end = k_off_h[L + 1];
            ++pl.n_lab;
            ++L;
        }
        pl.n_slots = last_end - k_off_h[L0];
        pl.label_stride = label_stride;
        const size_t smem = (size_t)pl.n_slots * kCellW * sizeof(long long) + 2 * kCellKeys * sizeof(unsigned int) +
                            2 * kCellRows * sizeof(unsigned short) + kCellKeys * sizeof(unsigned short);   // sacc | cursor, start | keys, order | segkey
        VTF_CUDA(cudaFuncSetAttribute(cell_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "vtf_cell_sums/smem");
        const unsigned ny = (unsigned)((d + kCellW - 1) / kCellW);
        VTF_REQUIRE(ny <= 16, "vtf_cell_sums: d <= 8192");
        int per_sm = 1;
        VTF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cell_sums_kernel, kCellThreads, smem), "vtf_cell_sums/occupancy");
        const long long chunks = (n + kCellRows - 1) / kCellRows;
        const unsigned gx = (unsigned)max(1ll, min(chunks, (long long)(num_sms() * max(per_sm, 1) / ny)));
        VTF_CUDA(cudaMemsetAsync(ws, 0, 16 * sizeof(unsigned int), st), "vtf_cell_sums/claim");
        VTF_LAUNCH(cell_sums_kernel, dim3(gx, ny), kCellThreads, smem, stream, X, (long long)n, d, labels, pl, (float)two_s,
                   reinterpret_cast<unsigned long long*>(acc + (size_t)k_off_h[L0] * d), bad, reinterpret_cast<unsigned int*>(ws));
        VTF_CHECK_LAST("vtf_cell_sums");
    }
    return VTF_OK;
}

}  // extern "C"

namespace vtf {

// ------------------------------------------------------------------------------------------------------------------
// cell_dispersion_kernel: per cluster, sum of |x - c|^2 and of |x - c| over its rows (KMeans.inertia_, calinski_harabasz's
// intra dispersion, davies_bouldin's distance sums; SK/cluster/_k_means_common.pyx:94-124, SK/metrics/cluster/
// _unsupervised.py:394-399, 467-474) for several labellings in one pass.  Same chunking and grouping as cell_sums_kernel;
// inside a cell every labelling measures against ONE centre, so a warp takes R rows of the cell at a time, keeps them in
// registers and reads each centre element from shared memory once per R rows.  The fp64 centre travels as two fp32 pieces
// (hi + lo, exact to 2^-48): e = (x - hi) - lo and the lane-strided sum of e^2 run in fp32 -- B200's CUDA cores retire fp64 at
// about 1/8 of the fp32 rate (measured: ~3e12 fp64 instructions/s in kpp_dist and in the kernel this replaces), which made
// an all-fp64 version of this kernel fp64-bound at 1.5 ms; the reference computes the inertia in fp32 too
// (_k_means_common.pyx:94-124).  Per row the squared distance is good to ~1e-6 relative; the per-row results are then
// accumulated as 64-bit FIXED-POINT integers -- exact, order-independent adds, so the result does not depend on the scatter
// order, the CTA schedule or the rank count.  The kernel this replaces (cluster_dispersion_kernel, accumulate.cu) re-read
// fp64 centre rows through L1 for every row and labelling: 0.16 of the HBM roofline for k = 2..6.
constexpr int kDispMaxLab = 12;      // labellings one dispersion launch measures (several may share a label row)

struct DispPlan {
    int n_key;                       // distinct label rows that make up the cell key
    int key_row[kCellMaxLab];        // row of the label matrix
    int key_kk[kCellMaxLab];         // its number of clusters
    int key_mult[kCellMaxLab];       // key = sum label * mult
    int n_lab;                       // labellings measured
    int lab_key[kDispMaxLab];        // which key component gives the labelling's label
    int slot0[kDispMaxLab];          // its first centre slot
    int n_slots;
    long long label_stride;
};

constexpr int kDispThreads = 512;    // 16 warps share one shared-memory copy of the centres (2 rows per warp at a time: <= 128 registers)
constexpr int kDispWarps = kDispThreads / 32;

template <int DV, int R>
__global__ void __launch_bounds__(kDispThreads, 1) cell_dispersion_kernel(const float* __restrict__ X, long long n, int d,
                                                                          const int32_t* __restrict__ labels, const DispPlan pl,
                                                                          const double* __restrict__ C64, double sq_unit, double rt_unit,
                                                                          unsigned long long* __restrict__ out, unsigned int* __restrict__ bad,
                                                                          unsigned int* __restrict__ claim) {
    extern __shared__ __align__(16) unsigned char cell_smem[];
    float* sc_hi = reinterpret_cast<float*>(cell_smem);                                       // [n_slots][d] centres: fp32(c)
    float* sc_lo = sc_hi + (size_t)pl.n_slots * d;                                            // [n_slots][d]          c - fp32(c)
    long long* wacc = reinterpret_cast<long long*>(sc_lo + (size_t)pl.n_slots * d);           // [16 warps][2][n_slots]
    unsigned int* cursor = reinterpret_cast<unsigned int*>(wacc + kDispWarps * 2 * pl.n_slots);
    unsigned int* start = cursor + kCellKeys;
    unsigned short* keys = reinterpret_cast<unsigned short*>(start + kCellKeys);
    unsigned short* order = keys + kCellRows;
    unsigned short* segkey = order + kCellRows;
    __shared__ unsigned int warp_tot[kDispWarps];
    __shared__ int s_nseg;
    __shared__ unsigned int s_chunk;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long n_chunks = (n + kCellRows - 1) / kCellRows;
    for (int i = t; i < pl.n_slots * d; i += kDispThreads) {
        const double c = C64[i];
        const float h = (float)c;
        sc_hi[i] = h;
        sc_lo[i] = (float)(c - (double)h);
    }
    for (int i = t; i < kDispWarps * 2 * pl.n_slots; i += kDispThreads) wacc[i] = 0ll;
    long long* my_sq = wacc + (size_t)warp * 2 * pl.n_slots;
    long long* my_rt = my_sq + pl.n_slots;
    unsigned int nbad = 0;
    for (;;) {
        if (t == 0) s_chunk = atomicAdd(claim, 1u);
        for (int i = t; i < kCellKeys; i += kDispThreads) cursor[i] = 0u;
        __syncthreads();
        if ((long long)s_chunk >= n_chunks) break;
        const long long row0 = (long long)s_chunk * kCellRows;
        const int Rn = (int)min((long long)kCellRows, n - row0);
        for (int i = t; i < Rn; i += kDispThreads) {
            int key = 0;
            for (int q = 0; q < pl.n_key; ++q) {
                int l = __ldg(labels + (long long)pl.key_row[q] * pl.label_stride + row0 + i);
                if ((unsigned)l >= (unsigned)pl.key_kk[q]) { ++nbad; l = 0; }
                key += l * pl.key_mult[q];
            }
            keys[i] = (unsigned short)key;
            atomicAdd(&cursor[key], 1u);
        }
        __syncthreads();
        constexpr int kBins = kCellKeys / kDispThreads;      // histogram bins per thread in the scan
        unsigned int c4[kBins];
        unsigned int local = 0;
#pragma unroll
        for (int j = 0; j < kBins; ++j) { c4[j] = cursor[kBins * t + j]; local += c4[j] + (c4[j] ? 0x10000u : 0u); }
        unsigned int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        unsigned int base = incl - local;
        for (int w = 0; w < warp; ++w) base += warp_tot[w];
        if (t == kDispThreads - 1) s_nseg = (int)((base + local) >> 16);
        {
            unsigned int pos = base & 0xffffu, seg = base >> 16;
#pragma unroll
            for (int j = 0; j < kBins; ++j) {
                start[kBins * t + j] = pos;
                cursor[kBins * t + j] = pos;
                if (c4[j]) segkey[seg++] = (unsigned short)(kBins * t + j);
                pos += c4[j];
            }
        }
        __syncthreads();
        for (int i = t; i < Rn; i += kDispThreads) order[atomicAdd(&cursor[keys[i]], 1u)] = (unsigned short)i;
        __syncthreads();

        const int nseg = s_nseg;
        for (int s = 0; s < nseg; ++s) {
            const int key = segkey[s];
            const int st = (int)start[key], en = (int)cursor[key];
            int slot[kDispMaxLab];                               // the cluster slot of this cell in every labelling
#pragma unroll
            for (int q = 0; q < kDispMaxLab; ++q)
                slot[q] = q < pl.n_lab ? pl.slot0[q] + (key / pl.key_mult[pl.lab_key[q]]) % pl.key_kk[pl.lab_key[q]] : 0;
            // the rows of the cell, R at a time per warp
            for (int j = st + R * warp; j < en; j += R * kDispWarps) {
                float4 xv[R][DV];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int jj = min(j + r, en - 1);
                    const float4* xr = reinterpret_cast<const float4*>(X + (size_t)(row0 + order[jj]) * d);
#pragma unroll
                    for (int v = 0; v < DV; ++v) xv[r][v] = __ldcs(xr + lane + 32 * v);
                }
                float acc[kDispMaxLab][R];
#pragma unroll
                for (int q = 0; q < kDispMaxLab; ++q)
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[q][r] = 0.f;
#pragma unroll
                for (int v = 0; v < DV; ++v) {
#pragma unroll
                    for (int q = 0; q < kDispMaxLab; ++q) {
                        if (q < pl.n_lab) {                      // uniform branch
                            const float4 ch = reinterpret_cast<const float4*>(sc_hi + (size_t)slot[q] * d)[lane + 32 * v];
                            const float4 cl = reinterpret_cast<const float4*>(sc_lo + (size_t)slot[q] * d)[lane + 32 * v];
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                float e;
                                e = (xv[r][v].x - ch.x) - cl.x; acc[q][r] = fmaf(e, e, acc[q][r]);
                                e = (xv[r][v].y - ch.y) - cl.y; acc[q][r] = fmaf(e, e, acc[q][r]);
                                e = (xv[r][v].z - ch.z) - cl.z; acc[q][r] = fmaf(e, e, acc[q][r]);
                                e = (xv[r][v].w - ch.w) - cl.w; acc[q][r] = fmaf(e, e, acc[q][r]);
                            }
                        }
                    }
                }
                // lane (q * R + r) % 32 takes the (labelling q, row r) result: square root and fixed-point conversion run on up
                // to 32 lanes at once instead of one after the other on lane 0 (two rounds when there are more than 32 results)
                float mine[2] = {0.f, 0.f};
                int my_slot[2] = {0, 0};
                bool have[2] = {false, false};
#pragma unroll
                for (int q = 0; q < kDispMaxLab; ++q) {
                    if (q < pl.n_lab) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const float sum = warp_sum(acc[q][r]);
                            constexpr int dummy = 0; (void)dummy;
                            const int idx = q * R + r;
                            if (lane == (idx & 31)) { mine[idx >> 5] = sum; my_slot[idx >> 5] = slot[q]; have[idx >> 5] = (j + r < en); }
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (have[h]) {
                        atomicAdd(reinterpret_cast<unsigned long long*>(my_sq + my_slot[h]), (unsigned long long)__double2ll_rn((double)mine[h] * sq_unit));
                        atomicAdd(reinterpret_cast<unsigned long long*>(my_rt + my_slot[h]), (unsigned long long)__double2ll_rn((double)sqrtf(mine[h]) * rt_unit));
                    }
                }
            }
        }
        __syncthreads();
    }
    if (nbad && bad) atomicAdd(bad, nbad);
    __syncthreads();
    for (int i = t; i < 2 * pl.n_slots; i += kDispThreads) {
        long long v = 0;
        for (int w = 0; w < kDispWarps; ++w) v += wacc[(size_t)w * 2 * pl.n_slots + i];
        if (v) atomicAdd(out + i, (unsigned long long)v);
    }
}

}  // namespace vtf

extern "C" {

// out[0 .. ktot) = sum |x - c|^2 * sq_unit, out[ktot .. 2 ktot) = sum |x - c| * rt_unit per centre slot, as int64 (zeroed here).
// Labelling q measures the rows against the centres C64[k_off[q] .. k_off[q+1]) using the labels of row row_of_h[q] of the
// label matrix (NULL: row q) -- several labellings may share a label row (a fit measured against its centres AND against the
// means of its final labels).  One launch: at most 12 labellings over at most 8 distinct label rows with at most 1024 joint
// label tuples, centres within 160 KB of shared memory; d % 128 == 0, d <= 1024.  The units must keep every total below
// 2^62: sq_unit <= 2^62 / (rows in total * max |x - c|^2), likewise rt_unit.
int vtf_cell_dispersion(const float* X, int64_t n, int d, const int32_t* labels, int64_t label_stride, int n_lab,
                        const int32_t* k_off_h, const int32_t* row_of_h, const double* C64, double sq_unit, double rt_unit,
                        int64_t* out, uint32_t* bad, void* ws, void* stream) {
    VTF_REQUIRE(X && labels && k_off_h && C64 && out && ws && n >= 0 && d > 0 && d % 128 == 0 && d <= 1024 && n_lab > 0 &&
                    n_lab <= kDispMaxLab && label_stride >= n && k_off_h[0] == 0 && sq_unit > 0.0 && rt_unit > 0.0,
                "vtf_cell_dispersion (d % 128 == 0, d <= 1024, at most 12 labellings)");
    cudaStream_t st = (cudaStream_t)stream;
    const int ktot = k_off_h[n_lab];
    VTF_CUDA(cudaMemsetAsync(out, 0, (size_t)2 * ktot * sizeof(int64_t), st), "vtf_cell_dispersion/zero");
    if (bad) VTF_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t), st), "vtf_cell_dispersion/zero");
    if (n == 0) return VTF_OK;
    DispPlan pl = {};
    long long keys = 1;
    for (int q = 0; q < n_lab; ++q) {
        const int k = k_off_h[q + 1] - k_off_h[q];
        const int row = row_of_h ? row_of_h[q] : q;
        VTF_REQUIRE(k > 0 && k <= kCellMaxSlots && row >= 0, "vtf_cell_dispersion: 1 <= k <= 48 per labelling");
        int kq = -1;
        for (int c = 0; c < pl.n_key; ++c) if (pl.key_row[c] == row) kq = c;
        if (kq < 0) {
            VTF_REQUIRE(pl.n_key < kCellMaxLab && keys * k <= kCellKeys, "vtf_cell_dispersion: more than 8 label rows or 1024 label tuples");
            kq = pl.n_key++;
            pl.key_row[kq] = row; pl.key_kk[kq] = k; pl.key_mult[kq] = (int)keys;
            keys *= k;
        }
        VTF_REQUIRE(pl.key_kk[kq] == k, "vtf_cell_dispersion: labellings that share a label row must have the same k");
        pl.lab_key[q] = kq;
        pl.slot0[q] = k_off_h[q];
    }
    pl.n_lab = n_lab;
    pl.n_slots = ktot;
    pl.label_stride = label_stride;
    VTF_REQUIRE((size_t)ktot * d * sizeof(double) <= 160 * 1024, "vtf_cell_dispersion: centres exceed 160 KB");
    const size_t smem = (size_t)ktot * d * sizeof(double) + (size_t)kDispWarps * 2 * ktot * sizeof(long long) + 2 * kCellKeys * sizeof(unsigned int) +
                        2 * kCellRows * sizeof(unsigned short) + kCellKeys * sizeof(unsigned short);
    unsigned int* claim = reinterpret_cast<unsigned int*>(ws);
    VTF_CUDA(cudaMemsetAsync(claim, 0, sizeof(unsigned int), st), "vtf_cell_dispersion/claim");
    const long long chunks = (n + kCellRows - 1) / kCellRows;
    const unsigned grid = (unsigned)max(1ll, min(chunks, (long long)num_sms()));
#define VTF_CD_CASE(DVv, Rv)                                                                                                        \
    case DVv:                                                                                                                       \
        VTF_CUDA(cudaFuncSetAttribute(cell_dispersion_kernel<DVv, Rv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),     \
                 "vtf_cell_dispersion/smem");                                                                                       \
        VTF_LAUNCH((cell_dispersion_kernel<DVv, Rv>), grid, kDispThreads, smem, stream, X, (long long)n, d, labels, pl, C64, sq_unit, \
                   rt_unit, reinterpret_cast<unsigned long long*>(out), bad, claim);                                                \
        break;
    switch (d / 128) {
        VTF_CD_CASE(1, 2) VTF_CD_CASE(2, 2) VTF_CD_CASE(3, 2) VTF_CD_CASE(4, 2)
        VTF_CD_CASE(5, 1) VTF_CD_CASE(6, 1) VTF_CD_CASE(7, 1) VTF_CD_CASE(8, 1)
    }
#undef VTF_CD_CASE
    VTF_CHECK_LAST("vtf_cell_dispersion");
    return VTF_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Cell layout of the silhouette sweep (gram_tc.cu): rows grouped by their joint label tuple over ALL labellings of the
// k-range, every cell padded to a multiple of 32 positions, the total to a multiple of 256.  Three launches, no sort:
// keys + histogram, one CTA that numbers the occupied cells (in key order) and lays them out, and the scatter of the row
// numbers.  Replaces the torch.unique / argsort / cumsum chain (two radix sorts, ~20 launches, three host syncs).
namespace vtf {

constexpr int kLayoutKeys = 4096;      // joint label tuples the layout kernels can tell apart

struct LayoutPlan {
    int n_lab;
    int kk[kMaxSeg];
    int mult[kMaxSeg];
    int n_keys;
    long long label_stride;
};

__global__ void __launch_bounds__(256) cell_keys_kernel(const int32_t* __restrict__ labels, long long N, const LayoutPlan pl,
                                                        unsigned short* __restrict__ keys, unsigned int* __restrict__ hist,
                                                        unsigned int* __restrict__ bad) {
    __shared__ unsigned int sh[kLayoutKeys];
    for (int i = threadIdx.x; i < pl.n_keys; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    unsigned int nbad = 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (long long)gridDim.x * blockDim.x) {
        int key = 0;
        for (int q = 0; q < pl.n_lab; ++q) {
            int l = __ldg(labels + (long long)q * pl.label_stride + r);
            if ((unsigned)l >= (unsigned)pl.kk[q]) { ++nbad; l = 0; }
            key += l * pl.mult[q];
        }
        keys[r] = (unsigned short)key;
        atomicAdd(&sh[key], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < pl.n_keys; i += blockDim.x) { const unsigned int v = sh[i]; if (v) atomicAdd(hist + i, v); }
    if (nbad) atomicAdd(bad, nbad);
}

// one CTA of 1024 threads, 4 keys per thread: dense cell ids in key order, padded offsets, chunk -> cell table,
// cell -> cluster tables, meta = {G, P, total}
__global__ void __launch_bounds__(1024) cell_layout_kernel(const unsigned int* __restrict__ hist, const LayoutPlan pl,
                                                           int32_t* __restrict__ cell_of_key, unsigned long long* __restrict__ cursor,
                                                           int32_t* __restrict__ chunk_cell, long long chunk_cap,
                                                           int32_t* __restrict__ c2c, int c2c_stride, long long* __restrict__ meta) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long s_off[kLayoutKeys + 1];     // padded start of dense cell g
    __shared__ unsigned long long s_total;
    __shared__ int s_G;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned int c4[4];
    unsigned long long local = 0;                                // low 40 bits: padded rows, high bits: occupied cells
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int key = 4 * t + j;
        c4[j] = key < pl.n_keys ? hist[key] : 0u;
        if (c4[j]) local += (unsigned long long)((c4[j] + 31u) & ~31u) + (1ull << 40);
    }
    unsigned long long incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned long long base = incl - local;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    if (t == 1023) { s_total = (base + local) & ((1ull << 40) - 1); s_G = (int)((base + local) >> 40); }
    {
        unsigned long long pos = base & ((1ull << 40) - 1);
        int g = (int)(base >> 40);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int key = 4 * t + j;
            if (key >= pl.n_keys) break;
            if (c4[j]) {
                cell_of_key[key] = g;
                cursor[g] = pos;
                s_off[g] = pos;
                for (int q = 0; q < pl.n_lab; ++q) c2c[(size_t)q * c2c_stride + g] = (key / pl.mult[q]) % pl.kk[q];
                pos += (c4[j] + 31u) & ~31u;
                ++g;
            } else {
                cell_of_key[key] = -1;
            }
        }
    }
    __syncthreads();
    const int G = s_G;
    const unsigned long long total = s_total;
    if (t == 0) {
        s_off[G] = total;
        meta[0] = G;
        meta[1] = (long long)((total + 255ull) & ~255ull);
        meta[2] = (long long)total;
    }
    __syncthreads();
    const long long n_chunks = min(chunk_cap, (long long)(((total + 255ull) & ~255ull) / 32ull));
    for (long long c = t; c < n_chunks; c += 1024) {
        const unsigned long long pos = (unsigned long long)c * 32ull;
        int g = 0;
        if (pos < total) {                                   // binary search: last g with s_off[g] <= pos
            int lo = 0, hi = G - 1;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_off[mid] <= pos) lo = mid; else hi = mid - 1; }
            g = lo;
        }
        chunk_cell[c] = g;                                   // chunks past the last cell (padding to 256 rows): cell 0, zero rows
    }
}

__global__ void __launch_bounds__(256) cell_scatter_kernel(const unsigned short* __restrict__ keys, long long N,
                                                           const int32_t* __restrict__ cell_of_key,
                                                           unsigned long long* __restrict__ cursor, int64_t* __restrict__ perm) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (long long)gridDim.x * blockDim.x) {
        // warp-aggregated claim: lanes of the same cell take consecutive positions with one atomic
        const int g = cell_of_key[keys[r]];
        const unsigned int peers = __match_any_sync(__activemask(), g);
        const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
        unsigned long long pos = 0;
        if (lane == leader) pos = atomicAdd(cursor + g, (unsigned long long)__popc(peers));
        pos = __shfl_sync(peers, pos, leader) + (unsigned long long)__popc(peers & ((1u << lane) - 1u));
        perm[pos] = r;
    }
}

}  // namespace vtf

extern "C" {

size_t vtf_build_cells_ws_bytes(int64_t N) {
    return (size_t)N * 2 + 16 + kLayoutKeys * (4 + 4 + 8) + 64;
}

int vtf_build_cells(const int32_t* labels, int64_t label_stride, int64_t N, int n_lab, const int32_t* k_h, int64_t* perm,
                    int64_t perm_cap, int32_t* chunk_cell, int32_t* c2c, int c2c_stride, int64_t* meta, void* ws, void* stream) {
    VTF_REQUIRE(labels && k_h && perm && chunk_cell && c2c && meta && ws && N > 0 && n_lab > 0 && n_lab <= kMaxSeg &&
                    label_stride >= N && perm_cap % 256 == 0,
                "vtf_build_cells");
    LayoutPlan pl = {};
    long long keys = 1;
    for (int q = 0; q < n_lab; ++q) {
        VTF_REQUIRE(k_h[q] > 0, "vtf_build_cells: k > 0");
        pl.kk[q] = k_h[q];
        pl.mult[q] = (int)keys;
        keys *= k_h[q];
        VTF_REQUIRE(keys <= kLayoutKeys, "vtf_build_cells: more than 4096 label tuples");
    }
    pl.n_lab = n_lab; pl.n_keys = (int)keys; pl.label_stride = label_stride;
    const long long g_max = min((long long)keys, (long long)N);
    VTF_REQUIRE(perm_cap >= ((N + 31 * g_max + 255) / 256) * 256 && c2c_stride >= g_max, "vtf_build_cells: perm_cap / c2c_stride too small");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* w = (unsigned char*)ws;
    unsigned int* hist = (unsigned int*)w;                               w += kLayoutKeys * 4;
    unsigned int* bad = (unsigned int*)w;                                w += 16;
    int32_t* cell_of_key = (int32_t*)w;                                  w += kLayoutKeys * 4;
    unsigned long long* cursor = (unsigned long long*)w;                 w += kLayoutKeys * 8;
    unsigned short* keybuf = (unsigned short*)w;
    VTF_CUDA(cudaMemsetAsync(hist, 0, kLayoutKeys * 4 + 16, st), "vtf_build_cells/zero");
    VTF_CUDA(cudaMemsetAsync(perm, 0xff, (size_t)perm_cap * sizeof(int64_t), st), "vtf_build_cells/perm");   // -1 = padding row
    const int grid = (int)min((int64_t)num_sms() * 8, (N + 255) / 256);
    VTF_LAUNCH(cell_keys_kernel, grid, 256, 0, stream, labels, (long long)N, pl, keybuf, hist, bad);
    VTF_CHECK_LAST("vtf_build_cells/keys");
    VTF_LAUNCH(cell_layout_kernel, 1, 1024, 0, stream, hist, pl, cell_of_key, cursor, chunk_cell, (long long)(perm_cap / 32), c2c,
               c2c_stride, (long long*)meta);
    VTF_CHECK_LAST("vtf_build_cells/layout");
    VTF_LAUNCH(cell_scatter_kernel, grid, 256, 0, stream, keybuf, (long long)N, cell_of_key, cursor, perm);
    VTF_CHECK_LAST("vtf_build_cells/scatter");
    return VTF_OK;
}

}  // extern "C"
