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

__device__ __forceinline__ float2 ldg_stream2(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

__global__ void __launch_bounds__(kCellThreads) cell_sums_kernel(const float* __restrict__ X, long long n, int d,
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
        if (j < en) {                                               // the tail re-reads the last row (L2 hit) and drops it
            float2 v[kCellUnroll];
#pragma unroll
            for (int u = 0; u < kCellUnroll; ++u)
                v[u] = ldg_stream2(reinterpret_cast<const float2*>(xcol + (size_t)order[min(j + u, en - 1)] * d));
#pragma unroll
            for (int u = 0; u < kCellUnroll; ++u) {
                if (j + u < en) {
                    a0 += __float2ll_rn(v[u].x * two_s);
                    a1 += __float2ll_rn(v[u].y * two_s);
                }
            }
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
