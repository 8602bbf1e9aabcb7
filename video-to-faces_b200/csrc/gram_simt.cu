// fp32 FMA Gram-tile kernels: the decision-exact paths.
//   cosine_dedup              <- dupes.py:60-65 (cosine_distances + lower-triangular min/argmin)
//   silhouette_rowsums_exact  <- SK/metrics/cluster/_unsupervised.py:149-200 (small N / cross-check)
// Both stream 64x64 tiles of X.X^T through shared memory and never materialise N x N (the
// reference allocates N*N fp32 plus an N*N fp64 np.tri temporary, dupes.py:60-62).
// Tensor-bound work done on the fp32 pipe on purpose: duplicate decisions and the logged
// distances must match an fp32 SGEMM, which TF32/fp16 operands do not (SURVEY.md section 7,
// hard part 2).  Algorithmic flops: N^2*d (dedup, lower triangle), 2*N^2*d (silhouette).
#include "common.cuh"

namespace vtf {

constexpr int kTile = 64;
constexpr int kKC = 16;

struct DedupEpi { unsigned long long* key; };
struct SilEpi { const float* sqn; const int32_t* labels; int k; double* rowsum; };

// MODE 0: dedup (only tiles with tj <= ti, columns j < i).  MODE 1: silhouette row sums.
template <int MODE>
__global__ void __launch_bounds__(256) gram_tile_kernel(const float* __restrict__ X, int64_t N, int d,
                                                        int64_t tile_i0, int64_t tile_istride,
                                                        DedupEpi de, SilEpi se) {
    const int64_t ti = tile_i0 + (int64_t)blockIdx.y * tile_istride;
    const int64_t tj = blockIdx.x;
    if (MODE == 0 && tj > ti) return;
    __shared__ __align__(16) float As[kKC][kTile + 4];
    __shared__ __align__(16) float Bs[kKC][kTile + 4];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;       // loader: row lrow, k offset lk
    const int64_t ai = min(ti * kTile + lrow, N - 1);    // clamp: out-of-range rows are masked later
    const int64_t bj = min(tj * kTile + lrow, N - 1);
    const float* ap = X + ai * d + lk;
    const float* bp = X + bj * d + lk;

    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

    float4 an = __ldg(reinterpret_cast<const float4*>(ap));
    float4 bn = __ldg(reinterpret_cast<const float4*>(bp));
    for (int k0 = 0; k0 < d; k0 += kKC) {
        As[lk + 0][lrow] = an.x; As[lk + 1][lrow] = an.y; As[lk + 2][lrow] = an.z; As[lk + 3][lrow] = an.w;
        Bs[lk + 0][lrow] = bn.x; Bs[lk + 1][lrow] = bn.y; Bs[lk + 2][lrow] = bn.z; Bs[lk + 3][lrow] = bn.w;
        __syncthreads();
        if (k0 + kKC < d) {
            an = __ldg(reinterpret_cast<const float4*>(ap + k0 + kKC));
            bn = __ldg(reinterpret_cast<const float4*>(bp + k0 + kKC));
        }
#pragma unroll
        for (int kk = 0; kk < kKC; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
        }
        __syncthreads();
    }

    const int64_t i0 = ti * kTile + ty * 4;
    const int64_t j0 = tj * kTile + tx * 4;
    if (MODE == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int64_t i = i0 + r;
            unsigned long long best = ~0ull;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int64_t j = j0 + c;
                if (i < N && j < i) {
                    const float dist = fminf(fmaxf(1.f - acc[r][c], 0.f), 2.f);   // pairwise.py:1177-1180
                    const unsigned long long key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)j;
                    best = key < best ? key : best;
                }
            }
            // the 16 threads that share row i sit in one half-warp
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other < best ? other : best;
            }
            if (tx == 0 && best != ~0ull) atomicMin(de.key + i, best);
        }
    } else {
        int lj[4];
        float nj[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int64_t j = j0 + c;
            lj[c] = j < N ? __ldg(se.labels + j) : -1;
            nj[c] = j < N ? __ldg(se.sqn + j) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int64_t i = i0 + r;
            if (i >= N) continue;
            const float ni = __ldg(se.sqn + i);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int64_t j = j0 + c;
                if (lj[c] < 0 || j == i) continue;
                const float d2 = fmaxf(fmaf(-2.f, acc[r][c], ni + nj[c]), 0.f);
                atomicAdd(se.rowsum + i * se.k + lj[c], (double)sqrtf(d2));
            }
        }
    }
}

__global__ void dedup_init_kernel(unsigned long long* key, int64_t N) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) key[i] = ((unsigned long long)__float_as_uint(10000.0f) << 32);   // dupes.py:62
}

}  // namespace vtf

using namespace vtf;

extern "C" {

int vtf_dedup_init(uint64_t* key, int64_t N, void* stream) {
    VTF_REQUIRE(key && N >= 0, "vtf_dedup_init");
    if (N == 0) return VTF_OK;
    VTF_LAUNCH(dedup_init_kernel, (unsigned)((N + 255) / 256), 256, 0, stream, (unsigned long long*)key, N);
    VTF_CHECK_LAST("vtf_dedup_init");
    return VTF_OK;
}

int vtf_cosine_dedup(const float* Xn, int64_t N, int d, int rank, int world, uint64_t* key, void* stream) {
    VTF_REQUIRE(Xn && key && N >= 0 && d > 0 && d % kKC == 0 && world > 0 && rank >= 0 && rank < world,
                "vtf_cosine_dedup (d must be a multiple of 16)");
    if (N <= 1) return VTF_OK;
    const int64_t nt = (N + kTile - 1) / kTile;
    if (rank >= nt) return VTF_OK;
    const int64_t mine = (nt - rank + world - 1) / world;   // tiles rank, rank+world, ...
    DedupEpi de{(unsigned long long*)key};
    SilEpi se{nullptr, nullptr, 0, nullptr};
    for (int64_t y0 = 0; y0 < mine; y0 += 32768) {
        const unsigned gy = (unsigned)min((int64_t)32768, mine - y0);
        VTF_LAUNCH(gram_tile_kernel<0>, dim3((unsigned)nt, gy), 256, 0, stream, Xn, N, d,
                   (int64_t)rank + y0 * world, (int64_t)world, de, se);
        VTF_CHECK_LAST("vtf_cosine_dedup");
    }
    return VTF_OK;
}

int vtf_silhouette_rowsums_exact(const float* X, const float* sqnorm, int64_t N, int d, const int32_t* labels, int k,
                                 int64_t row_begin, int64_t row_end, double* rowsum, void* stream) {
    VTF_REQUIRE(X && sqnorm && labels && rowsum && N >= 0 && d > 0 && d % kKC == 0 && k > 0 && row_begin >= 0 &&
                    row_end <= N && row_begin % kTile == 0,
                "vtf_silhouette_rowsums_exact (d % 16 == 0, row_begin % 64 == 0)");
    if (row_end <= row_begin) return VTF_OK;
    const int64_t nt = (N + kTile - 1) / kTile;
    const int64_t t0 = row_begin / kTile, t1 = (row_end + kTile - 1) / kTile;
    DedupEpi de{nullptr};
    SilEpi se{sqnorm, labels, k, rowsum};
    for (int64_t y0 = t0; y0 < t1; y0 += 32768) {
        const unsigned gy = (unsigned)min((int64_t)32768, t1 - y0);
        VTF_LAUNCH(gram_tile_kernel<1>, dim3((unsigned)nt, gy), 256, 0, stream, X, N, d, y0, (int64_t)1, de, se);
        VTF_CHECK_LAST("vtf_silhouette_rowsums_exact");
    }
    return VTF_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// 'hash' branch of remove_dupes_overall (dupes.py:55-58, called from detection.py:57-59): all-pairs
// Hamming distance between 64-bit average hashes (dupes.py:11-15), each row against every EARLIER row.
// The reference evaluates a Python callable N^2 times (SK/metrics/pairwise.py:2034-2051); here one
// thread owns a row, hashes stream through shared memory, distance = popc(xor).
namespace vtf {

// warp per row: bits[N,64] int64 (0/1, the layout ahash() returns after np.stack) -> one uint64 per row
__global__ void __launch_bounds__(256) hash_pack_kernel(const long long* __restrict__ bits, int64_t N,
                                                        unsigned long long* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * 8;
    for (; row < N; row += stride) {
        const unsigned lo = __ballot_sync(0xffffffffu, bits[row * 64 + lane] != 0);
        const unsigned hi = __ballot_sync(0xffffffffu, bits[row * 64 + 32 + lane] != 0);
        if (lane == 0) out[row] = ((unsigned long long)hi << 32) | lo;
    }
}

constexpr int kHashRows = 256, kHashCols = 1024;

__global__ void __launch_bounds__(kHashRows) hash_dedup_kernel(const unsigned long long* __restrict__ h, int64_t N,
                                                               int64_t tile0, int64_t tile_stride,
                                                               unsigned long long* __restrict__ key) {
    __shared__ unsigned long long sh[kHashCols];
    const int64_t ti = tile0 + (int64_t)blockIdx.x * tile_stride;
    const int64_t i = ti * kHashRows + threadIdx.x;
    const unsigned long long hi = i < N ? h[i] : 0ull;
    unsigned long long best = ~0ull;
    const int64_t jend = min(N, (ti + 1) * kHashRows);        // columns j < i only
    for (int64_t j0 = 0; j0 < jend; j0 += kHashCols) {
        __syncthreads();
        for (int t = threadIdx.x; t < kHashCols; t += kHashRows) sh[t] = (j0 + t < N) ? h[j0 + t] : 0ull;
        __syncthreads();
        const int lim = (int)min((int64_t)kHashCols, i - j0);   // j < i
        for (int t = 0; t < lim; ++t) {
            const unsigned long long k = ((unsigned long long)__popcll(hi ^ sh[t]) << 32) | (unsigned long long)(j0 + t);
            best = k < best ? k : best;                         // ascending j: first index wins ties
        }
    }
    if (i < N && i > 0) key[i] = best;
}

__global__ void hash_init_kernel(unsigned long long* key, int64_t N) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) key[i] = 10000ull << 32;                         // dupes.py:62 in uint16 arithmetic
}

}  // namespace vtf

extern "C" {

int vtf_hash_pack(const int64_t* bits, int64_t N, uint64_t* packed, void* stream) {
    VTF_REQUIRE(bits && packed && N >= 0, "vtf_hash_pack");
    if (N == 0) return VTF_OK;
    const int grid = (int)min((int64_t)num_sms() * 8, (N + 7) / 8);
    VTF_LAUNCH(hash_pack_kernel, grid, 256, 0, stream, (const long long*)bits, N, (unsigned long long*)packed);
    VTF_CHECK_LAST("vtf_hash_pack");
    return VTF_OK;
}

int vtf_hash_dedup(const uint64_t* packed, int64_t N, int rank, int world, uint64_t* key, void* stream) {
    VTF_REQUIRE(packed && key && N >= 0 && world > 0 && rank >= 0 && rank < world, "vtf_hash_dedup");
    if (N == 0) return VTF_OK;
    VTF_LAUNCH(hash_init_kernel, (unsigned)((N + 255) / 256), 256, 0, stream, (unsigned long long*)key, N);
    VTF_CHECK_LAST("vtf_hash_dedup/init");
    const int64_t nt = (N + kHashRows - 1) / kHashRows;
    if (rank >= nt) return VTF_OK;
    const int64_t mine = (nt - rank + world - 1) / world;
    VTF_LAUNCH(hash_dedup_kernel, (unsigned)mine, kHashRows, 0, stream, (const unsigned long long*)packed, N,
               (int64_t)rank, (int64_t)world, (unsigned long long*)key);
    VTF_CHECK_LAST("vtf_hash_dedup");
    return VTF_OK;
}

}  // extern "C"
