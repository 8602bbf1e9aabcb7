// k-means++ seeding kernels (SK/cluster/_kmeans.py:180-278).
//   kpp_dist: distances from every row to T candidate centres, computed in fp64 and stored as
//             fp32 exactly like sklearn's upcast path (SK/metrics/pairwise.py:567-638), fused
//             with min(closest, .) (:262) and the per-candidate potential (:263).
//             HBM-bound: N*d*4 bytes read, T*N*4 written.
//   kpp_pick: searchsorted over the SEQUENTIAL fp32 cumsum of the closest distances (:250).
//             The running sum is one dependent chain of N fp32 additions, so one thread walks
//             it while the rest of the CTA streams the next chunk into shared memory; a
//             parallel scan would round differently and move candidates by a few rows.
#include <limits.h>
#include <stdlib.h>
#include "common.cuh"

namespace vtf {

constexpr int kMaxTrials = 8;

// |x_row|^2 in fp64, in exactly the order kpp_dist_kernel accumulates it (lane-strided float4, four FMAs each, butterfly
// warp_sum): computed once per fit context and handed to every seeding round instead of being recomputed 8 to 61 times.
__global__ void __launch_bounds__(256) kpp_rownorm64_kernel(const float* __restrict__ X, int64_t N, int d, double* __restrict__ xsq) {
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * 8;
    for (; row < N; row += stride) {
        const float4* xr = reinterpret_cast<const float4*>(X + row * d);
        double xx = 0.0;
        for (int v = lane; v < nvec; v += 32) {
            const float4 f = ldg_stream(xr + v);
            const double a0 = f.x, a1 = f.y, a2 = f.z, a3 = f.w;
            xx = fma(a0, a0, xx); xx = fma(a1, a1, xx); xx = fma(a2, a2, xx); xx = fma(a3, a3, xx);
        }
        xx = warp_sum(xx);
        if (lane == 0) xsq[row] = xx;
    }
}

template <int TT, int R, bool HAVE_XSQ>   // TT = compile-time bound on the number of candidates, R = rows per warp and pass
__global__ void __launch_bounds__(256, 2) kpp_dist_kernel(const float* __restrict__ X, int64_t N, int d,
                                                       const float* __restrict__ cand, int T,
                                                       const float* __restrict__ closest, const double* __restrict__ xsq,
                                                       float* __restrict__ out, double* __restrict__ part) {
    extern __shared__ __align__(16) double sc[];   // [T][d] candidates, converted to fp64 once per CTA
    __shared__ double s_cc[kMaxTrials];
    __shared__ double s_pot[8][kMaxTrials];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = d >> 2;
    for (int i = threadIdx.x; i < T * d; i += blockDim.x) sc[i] = (double)__ldg(cand + i);
    __syncthreads();
    if (warp < T) {
        double s = 0.0;
        for (int j = lane; j < d; j += 32) { const double c = sc[warp * d + j]; s = fma(c, c, s); }
        s = warp_sum(s);
        if (lane == 0) s_cc[warp] = s;
    }
    __syncthreads();

    double pot[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) pot[t] = 0.0;

    // R rows per warp and pass (4 for up to 4 candidates, 2 above: registers): every candidate element read from shared memory serves R
    // rows -- the kernel was bound by those reads (T wavefronts per row and float4), not by HBM or the fp64 pipe.
    // Every row keeps its own accumulation order, so the results are bit-identical to the one-row form.
    int64_t row0 = ((int64_t)blockIdx.x * 8 + warp) * R;
    const int64_t stride = (int64_t)gridDim.x * 8 * R;
    for (; row0 < N; row0 += stride) {
        const float4* xr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) xr[r] = reinterpret_cast<const float4*>(X + (row0 + r < N ? row0 + r : row0) * d);
        double xx[R], dot[R][TT];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            xx[r] = 0.0;
#pragma unroll
            for (int t = 0; t < TT; ++t) dot[r][t] = 0.0;
        }
        for (int v = lane; v < nvec; v += 32) {
            double a[R][4];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 f = ldg_stream(xr[r] + v);
                a[r][0] = f.x; a[r][1] = f.y; a[r][2] = f.z; a[r][3] = f.w;
                if (!HAVE_XSQ) {
                    xx[r] = fma(a[r][0], a[r][0], xx[r]); xx[r] = fma(a[r][1], a[r][1], xx[r]);
                    xx[r] = fma(a[r][2], a[r][2], xx[r]); xx[r] = fma(a[r][3], a[r][3], xx[r]);
                }
            }
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                if (t < T) {
                    const double2 c01 = reinterpret_cast<const double2*>(sc + (size_t)t * d)[2 * v];
                    const double2 c23 = reinterpret_cast<const double2*>(sc + (size_t)t * d)[2 * v + 1];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        dot[r][t] = fma(a[r][0], c01.x, dot[r][t]); dot[r][t] = fma(a[r][1], c01.y, dot[r][t]);
                        dot[r][t] = fma(a[r][2], c23.x, dot[r][t]); dot[r][t] = fma(a[r][3], c23.y, dot[r][t]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = row0 + r;
            const bool have = row < N;                                    // warp-uniform
            const double xs = HAVE_XSQ ? (have ? __ldg(xsq + row) : 0.0) : warp_sum(xx[r]);
            const float prev = (closest && have) ? __ldg(closest + row) : INFINITY;
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                if (t < T) {
                    const double dt = warp_sum(dot[r][t]);
                    // ((-2 x.c) + |c|^2) + |x|^2 in fp64, then fp32, then clamp (pairwise.py:632-636, 413)
                    float dist = (float)((-2.0 * dt + s_cc[t]) + xs);
                    dist = fmaxf(dist, 0.f);
                    dist = fminf(dist, prev);
                    if (lane == 0 && have) {
                        out[(size_t)t * N + row] = dist;
                        pot[t] += (double)dist;
                    }
                }
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < TT; ++t) s_pot[warp][t] = pot[t];
    }
    __syncthreads();
    if (threadIdx.x < T) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += s_pot[w][threadIdx.x];
        part[(size_t)blockIdx.x * T + threadIdx.x] = s;
    }
}

constexpr int kPickChunk = 4096;

__global__ void __launch_bounds__(128) kpp_pick_kernel(const float* __restrict__ closest, int64_t N,
                                                       const float* __restrict__ thr, int T,
                                                       const float* __restrict__ offset_in, float* __restrict__ total_out,
                                                       int64_t* __restrict__ idx) {
    __shared__ __align__(16) float buf[2][kPickChunk];
    __shared__ float sthr[kMaxTrials];
    if (threadIdx.x < T) { sthr[threadIdx.x] = thr[threadIdx.x]; idx[threadIdx.x] = -1; }
    const int64_t nchunks = (N + kPickChunk - 1) / kPickChunk;
    auto load = [&](int64_t c) {
        float* b = buf[c & 1];
        const int64_t base = c * kPickChunk;
        for (int i = threadIdx.x; i < kPickChunk; i += blockDim.x) {
            const int64_t g = base + i;
            b[i] = g < N ? __ldg(closest + g) : 0.f;
        }
    };
    if (nchunks > 0) load(0);
    __syncthreads();
    float s = offset_in ? *offset_in : 0.f;
    int t = 0;
    // thresholds already met by the incoming running sum belong to an earlier shard
    if (threadIdx.x == 0) { while (t < T && s >= sthr[t]) ++t; }
    float th = (threadIdx.x == 0 && t < T) ? sthr[t] : INFINITY;
    for (int64_t c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) load(c + 1);
        if (threadIdx.x == 0) {
            // Distances are >= 0, so the running sum never decreases: test one prefix per 16 elements
            // and replay the block element by element only when a threshold falls inside it.  The
            // chunk is zero-padded past N (s + 0 == s), so whole blocks are always safe.
            const float4* b4 = reinterpret_cast<const float4*>(buf[c & 1]);
            const int64_t base = c * kPickChunk;
            // software pipeline: the 4 x LDS.128 of block i+1 are in flight while the 16-add chain of
            // block i (the critical path: 16 x 4 cycles) runs
            float4 nq[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) nq[u] = b4[u];
#pragma unroll 1
            for (int i = 0; i < kPickChunk / 4; i += 8) {      // 32 elements per block: one compare per 32 adds
                float4 q[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) q[u] = nq[u];
                if (i + 8 < kPickChunk / 4) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) nq[u] = b4[i + 8 + u];
                }
                float e = s;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    e = __fadd_rn(e, q[u].x); e = __fadd_rn(e, q[u].y); e = __fadd_rn(e, q[u].z); e = __fadd_rn(e, q[u].w);
                }
                if (e >= th) {
                    float r = s;
                    const float* bs = buf[c & 1] + 4 * i;   // rare: replay the block element by element
#pragma unroll 1
                    for (int j = 0; j < 32; ++j) {
                        r = __fadd_rn(r, bs[j]);
                        while (r >= th) {      // searchsorted side='left': first i with cumsum[i] >= v
                            const int64_t g = base + 4 * i + j;
                            idx[t] = g < N ? g : -1;
                            ++t;
                            th = t < T ? sthr[t] : INFINITY;
                        }
                    }
                }
                s = e;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = s;
}

// ------------------------------------------------------------------------------------------------
// kpp_pick_scan_kernel: the same searchsorted over the sequential fp32 cumsum, BIT-EXACT, but parallel.
//
// While the running sum S stays inside one binade [2^e, 2^(e+1)) it is an integer multiple q of u = 2^(e-23),
// and fl(S + v) = u * (q + a + r) with a = floor(v/u), r = round-to-nearest of the fraction f = v/u - a.
// Except for exact ties (f == 1/2, where round-half-even looks at the parity of q + a) the increment a + r does
// not depend on S, so a block-wide integer prefix sum reproduces every intermediate value of the sequential
// chain.  The chain is only walked one element at a time at "events": a tie, an addend >= 2^e, the element
// that carries S into the next binade, and S == 0.  Between events one CTA scans a shared-memory tile of
// 8192 elements per pass.  About 30 events per million rows: ~0.3 ms instead of the ~4 ms of the plain chain.
constexpr int kScanThreads = 1024;
constexpr int kScanPer = 8;
constexpr int kScanTile = kScanThreads * kScanPer;

// FAST: the tiles are not all walked.  Two multi-CTA passes (kpp_chunk_sum_kernel, kpp_chunk_int_kernel) have computed, for
// every tile j, the binade e_j its incoming running sum is expected to lie in and the exact integer total C_j of the
// rounded addends under that binade (or "unsafe": a tie, an oversized addend, a zero running sum).  Thread 0 then jumps
// over every tile for which the TRUE running sum S (known exactly, sequentially) is in binade e_j, S + C_j u stays inside
// the binade and no pending threshold is reached: S' = S + C_j u is then exactly what the chain of 8192 fp32 additions
// gives.  Only the other tiles -- binade crossings (~log2 of the tile count), ties, the tiles that hold a threshold --
// are walked by the block scan below.  Same results bit for bit; ~15 of 122 tiles walked at 1M rows.
template <bool FAST>
__global__ void __launch_bounds__(kScanThreads) kpp_pick_scan_kernel(const float* __restrict__ closest, int64_t N,
                                                                     const float* __restrict__ thr, int T,
                                                                     const float* __restrict__ offset_in,
                                                                     float* __restrict__ total_out, int64_t* __restrict__ idx,
                                                                     const long long* __restrict__ chunk_c,
                                                                     const int* __restrict__ chunk_e) {
    __shared__ float tile[kScanTile];
    __shared__ long long s_next;           // FAST: first tile that has to be walked
    __shared__ long long warp_tot[32];
    __shared__ int s_ev, s_evthr;          // earliest restart event / earliest threshold hit (tile-relative)
    __shared__ float s_sbefore;            // running sum just before the restart event
    __shared__ float s_S;
    __shared__ int s_t;
    __shared__ float sthr[kMaxTrials];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < T) { sthr[tid] = thr[tid]; idx[tid] = -1; }
    if (tid == 0) {
        float S0 = offset_in ? *offset_in : 0.f;
        int t0 = 0;
        while (t0 < T && S0 >= thr[t0]) ++t0;     // met by an earlier shard
        s_S = S0; s_t = t0;
    }
    __syncthreads();
    // the next tile's elements wait in registers while the current one is scanned (one CTA cannot hide HBM latency otherwise)
    float nxt[kScanPer];
    if (!FAST) {
#pragma unroll
        for (int k = 0; k < kScanPer; ++k) { const int64_t g = (int64_t)k * kScanThreads + tid; nxt[k] = g < N ? __ldg(closest + g) : 0.f; }
    }
    for (int64_t base = 0; base < N; base += kScanTile) {
        if (s_t >= T && total_out == nullptr) break;                   // every threshold placed (uniform: s_t is shared)
        if (FAST) {
            if (tid == 0) {
                const long long n_tiles = (long long)((N + kScanTile - 1) / kScanTile);
                long long j = (long long)(base / kScanTile);
                float S = s_S;
                const int t = s_t;
                while (j < n_tiles) {
                    if (!(S > 0.f)) break;
                    const int e = (int)((__float_as_uint(S) >> 23) & 0xff) - 127;
                    if (e < -100 || chunk_e[j] != e) break;
                    const long long q0 = (long long)(__float_as_uint(S) & 0x7fffff) + (1ll << 23);
                    const long long qn = q0 + chunk_c[j];
                    if (qn >= (1ll << 24)) break;                              // the tile carries S into the next binade
                    if (t < T) {
                        const double cq = ceil((double)sthr[t] * (double)__uint_as_float((unsigned)(23 - e + 127) << 23));
                        if (cq < 9.0e18 && qn >= (long long)cq) break;         // the pending threshold is reached inside this tile
                    }
                    S = (float)qn * __uint_as_float((unsigned)(e - 23 + 127) << 23);     // exact: qn < 2^24
                    ++j;
                }
                s_S = S;
                s_next = j;
            }
            __syncthreads();
            base = (int64_t)s_next * kScanTile;
            if (base >= N) break;
        }
        const int n = (int)min((int64_t)kScanTile, N - base);
        if (FAST) {
#pragma unroll
            for (int k = 0; k < kScanPer; ++k) {
                const int64_t g = base + (int64_t)k * kScanThreads + tid;
                tile[k * kScanThreads + tid] = g < N ? __ldg(closest + g) : 0.f;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kScanPer; ++k) tile[k * kScanThreads + tid] = nxt[k];
#pragma unroll
            for (int k = 0; k < kScanPer; ++k) {
                const int64_t g = base + kScanTile + (int64_t)k * kScanThreads + tid;
                nxt[k] = g < N ? __ldg(closest + g) : 0.f;
            }
        }
        __syncthreads();
        int pos = 0;
        while (pos < n) {
            const float S = s_S;
            const int t = s_t;
            if (t >= T && total_out == nullptr) break;                  // nothing left to find
            // --- per-thread pass over its kScanPer consecutive elements of [pos, n)
            const int rem = n - pos;
            const int per = (rem + kScanThreads - 1) / kScanThreads;   // <= kScanPer
            const int my0 = pos + tid * per;
            const int my1 = min(n, my0 + per);
            const bool Szero = !(S > 0.f);
            const int e = Szero ? 0 : (int)((__float_as_uint(S) >> 23) & 0xff) - 127;
            const bool slow = Szero || e < -100;                       // tiny sums: walk them
            const float two_e = slow ? 0.f : __uint_as_float((unsigned)(e + 127) << 23);
            const float scale = slow ? 0.f : __uint_as_float((unsigned)(23 - e + 127) << 23);   // 1/u = 2^(23-e)
            long long c[kScanPer];
            long long local = 0;
            int first_ev = INT_MAX;
#pragma unroll
            for (int k = 0; k < kScanPer; ++k) {
                c[k] = 0;
                const int i = my0 + k;
                if (k < per && i < my1) {
                    const float v = tile[i];
                    if (v != 0.f) {
                        if (slow || v >= two_e) { first_ev = min(first_ev, i); }
                        else {
                            const float tt = v * scale;                 // exact: power-of-two scaling, tt < 2^23
                            const int a = (int)tt;
                            const float f = tt - (float)a;
                            if (f == 0.5f) first_ev = min(first_ev, i);
                            else c[k] = (long long)a + (f > 0.5f ? 1 : 0);
                        }
                    }
                    local += c[k];
                }
            }
            // --- block-wide exclusive prefix of `local`
            long long incl = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (lane == 31) warp_tot[warp] = incl;
            if (tid == 0) { s_ev = INT_MAX; s_evthr = INT_MAX; }
            __syncthreads();
            if (warp == 0) {
                long long w = warp_tot[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long up = __shfl_up_sync(0xffffffffu, w, o);
                    if (lane >= o) w += up;
                }
                warp_tot[lane] = w;                                     // inclusive over warps
            }
            __syncthreads();
            const long long excl = (warp ? warp_tot[warp - 1] : 0) + (incl - local);
            const long long total = warp_tot[31];
            // --- walk my elements again with the true q: binade crossing and threshold hit
            const long long q0 = slow ? 0 : (long long)(__float_as_uint(S) & 0x7fffff) + (1ll << 23);   // S = q0 * u
            long long need = LLONG_MAX;
            if (!slow && t < T) {
                const double tq = (double)sthr[t] * (double)scale;       // exact
                const double cq = ceil(tq);
                need = cq >= 9.0e18 ? LLONG_MAX : (long long)cq;
            }
            long long q = q0 + excl;
            int my_ev = first_ev, my_thr = INT_MAX;
            float sb = 0.f;
            bool have_sb = false;
#pragma unroll
            for (int k = 0; k < kScanPer; ++k) {
                const int i = my0 + k;
                if (k < per && i < my1 && i <= my_ev) {
                    if (i == my_ev) { sb = (float)q * (slow ? 0.f : 1.f) ; have_sb = true; }
                    else {
                        const long long qn = q + c[k];
                        if (!slow && qn >= (1ll << 24)) { my_ev = i; sb = (float)q; have_sb = true; }
                        else {
                            if (qn >= need && my_thr == INT_MAX) my_thr = i;
                            q = qn;
                        }
                    }
                }
            }
            if (my_ev != INT_MAX) atomicMin(&s_ev, my_ev);
            if (my_thr != INT_MAX) atomicMin(&s_evthr, my_thr);
            __syncthreads();
            const int ev = s_ev, evthr = s_evthr;
            if (have_sb && my_ev == ev) s_sbefore = slow ? S : sb * __uint_as_float((unsigned)(e - 23 + 127) << 23);   // q * u, exact
            __syncthreads();
            if (evthr < ev) {
                // the current threshold is first reached at evthr, before anything that needs the serial chain
                if (tid == 0) { idx[t] = base + evthr; s_t = t + 1; }
                __syncthreads();
                continue;                                               // same pos, same S, next threshold
            }
            if (ev == INT_MAX) {
                if (tid == 0) s_S = slow ? S : (float)(q0 + total) * __uint_as_float((unsigned)(e - 23 + 127) << 23);
                pos = n;
                __syncthreads();
                continue;
            }
            // serial step at the event element, then restart right after it
            if (tid == 0) {
                float Sn = __fadd_rn(s_sbefore, tile[ev]);
                int tt = t;
                while (tt < T && Sn >= sthr[tt]) { idx[tt] = base + ev; ++tt; }
                s_S = Sn; s_t = tt;
            }
            pos = ev + 1;
            __syncthreads();
        }
        __syncthreads();
    }
    if (tid == 0 && total_out) *total_out = s_S;
}

// fp64 sum of every tile of kScanTile elements (approximate running sums at the tile boundaries)
__global__ void __launch_bounds__(256) kpp_chunk_sum_kernel(const float* __restrict__ closest, int64_t N, double* __restrict__ csum) {
    __shared__ double sred[8];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    double s = 0.0;
    for (int i = threadIdx.x; i < kScanTile; i += 256) { const int64_t g = base + i; if (g < N) s += (double)__ldg(closest + g); }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += sred[w]; csum[blockIdx.x] = t; }
}

// Tile j under the binade its incoming running sum is EXPECTED to lie in (from the fp64 tile sums): C_j = sum of the rounded
// addends a + r (see kpp_pick_scan_kernel), chunk_e[j] = that binade -- or INT_MIN when the tile holds an event (a tie
// f == 1/2, an addend >= 2^e) or the expected sum is zero / tiny, so that the walk does not skip it.
__global__ void __launch_bounds__(256) kpp_chunk_int_kernel(const float* __restrict__ closest, int64_t N, const double* __restrict__ csum,
                                                            const float* __restrict__ offset_in, long long* __restrict__ chunk_c,
                                                            int* __restrict__ chunk_e) {
    __shared__ double sred[8];
    __shared__ long long ired[8];
    __shared__ float s_S;
    const int j = blockIdx.x;
    double pre = 0.0;
    for (int i = threadIdx.x; i < j; i += 256) pre += csum[i];
    pre = warp_sum(pre);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = pre;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = offset_in ? (double)*offset_in : 0.0;
        for (int w = 0; w < 8; ++w) t += sred[w];
        s_S = (float)t;
    }
    __syncthreads();
    const float S = s_S;
    const int e = (S > 0.f) ? (int)((__float_as_uint(S) >> 23) & 0xff) - 127 : -1000;
    if (e < -100) { if (threadIdx.x == 0) { chunk_c[j] = 0; chunk_e[j] = INT_MIN; } return; }
    const float two_e = __uint_as_float((unsigned)(e + 127) << 23);
    const float scale = __uint_as_float((unsigned)(23 - e + 127) << 23);
    const int64_t base = (int64_t)j * kScanTile;
    long long c = 0;
    int ev = 0;
    for (int i = threadIdx.x; i < kScanTile; i += 256) {
        const int64_t g = base + i;
        const float v = g < N ? __ldg(closest + g) : 0.f;
        if (v != 0.f) {
            if (!(v < two_e)) ev = 1;                       // oversized addend (or NaN): walk it
            else {
                const float tt = v * scale;                 // exact: power-of-two scaling, tt < 2^23
                const int a = (int)tt;
                const float f = tt - (float)a;
                if (f == 0.5f) ev = 1;                      // round-half-even looks at the running sum
                else c += (long long)a + (f > 0.5f ? 1 : 0);
            }
        }
    }
    ev = __syncthreads_or(ev);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) ired[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < 8; ++w) t += ired[w];
        chunk_c[j] = t;
        chunk_e[j] = ev ? INT_MIN : e;
    }
}

// The same distances on the fp64 TENSOR pipe.  The CUDA cores of B200 retire DFMA at about 1/8 of the fp32 rate, which made
// the kernel above fp64-bound (0.45 of HBM at T = 3); mma.sync.m8n8k4.f64 computes an 8 x 8 x 4 block of exact products with
// fp64 accumulation per instruction, and the candidates (T <= 8) fill its N dimension.  Warp = 16 rows per pass (two 8-row
// groups that share every candidate read); lane (r = lane / 4, q = lane % 4) loads the float4 at columns 16 j + 4 q of row r
// and feeds element m of it to MMA m of the step, the B operand taking the matching candidate elements from shared memory
// (fp64, rows padded by 16 B so that the eight candidates start in different banks).  Products of two fp32 values are exact in
// fp64, so only the ORDER of the fp64 additions differs from the kernel above (and from any BLAS): after the rounding to fp32
// the distances are equal except where the fp64 sums straddle a rounding boundary (test_kpp_dist_matches_upcast_path).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 3) kpp_dist_mma_kernel(const float* __restrict__ X, int64_t N, int d, const float* __restrict__ cand,
                                                           int T, const float* __restrict__ closest, const double* __restrict__ xsq,
                                                           float* __restrict__ out, double* __restrict__ part) {
    extern __shared__ __align__(16) double sc[];   // [T][d + 2] candidates in fp64
    __shared__ double s_cc[kMaxTrials];
    __shared__ double s_pot[8][kMaxTrials];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ld = d + 2;
    for (int i = threadIdx.x; i < T * d; i += blockDim.x) sc[(i / d) * ld + (i % d)] = (double)__ldg(cand + i);
    __syncthreads();
    if (warp < T) {
        double s = 0.0;
        for (int j = lane; j < d; j += 32) { const double c = sc[warp * ld + j]; s = fma(c, c, s); }
        s = warp_sum(s);
        if (lane == 0) s_cc[warp] = s;
    }
    __syncthreads();

    const int r = lane >> 2, q = lane & 3;
    const bool bcol = r < T;                                   // this lane's B column is a real candidate
    const double* bsrc = sc + (size_t)(bcol ? r : 0) * ld + 4 * q;
    const int t0 = 2 * q, t1 = 2 * q + 1;                      // the candidates whose results land in this lane
    double pot0 = 0.0, pot1 = 0.0;
    const int nchunk = d >> 4;                                 // 16 columns per step

    int64_t row0 = ((int64_t)blockIdx.x * 8 + warp) * 16;
    const int64_t stride = (int64_t)gridDim.x * 8 * 16;
    for (; row0 < N; row0 += stride) {
        const int64_t ra = row0 + r, rb = row0 + 8 + r;
        const float4* xa = reinterpret_cast<const float4*>(X + (ra < N ? ra : row0) * d) + q;
        const float4* xb = reinterpret_cast<const float4*>(X + (rb < N ? rb : row0) * d) + q;
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        int j = 0;
        for (; j + 4 <= nchunk; j += 4) {
            float4 fa[4], fb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { fa[u] = ldg_stream(xa + 4 * (j + u)); fb[u] = ldg_stream(xb + 4 * (j + u)); }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                double2 c01 = make_double2(0.0, 0.0), c23 = c01;
                if (bcol) {
                    c01 = *reinterpret_cast<const double2*>(bsrc + 16 * (j + u));
                    c23 = *reinterpret_cast<const double2*>(bsrc + 16 * (j + u) + 2);
                }
                dmma884(a0, a1, (double)fa[u].x, c01.x); dmma884(b0, b1, (double)fb[u].x, c01.x);
                dmma884(a0, a1, (double)fa[u].y, c01.y); dmma884(b0, b1, (double)fb[u].y, c01.y);
                dmma884(a0, a1, (double)fa[u].z, c23.x); dmma884(b0, b1, (double)fb[u].z, c23.x);
                dmma884(a0, a1, (double)fa[u].w, c23.y); dmma884(b0, b1, (double)fb[u].w, c23.y);
            }
        }
        for (; j < nchunk; ++j) {
            const float4 fa = ldg_stream(xa + 4 * j), fb = ldg_stream(xb + 4 * j);
            double2 c01 = make_double2(0.0, 0.0), c23 = c01;
            if (bcol) {
                c01 = *reinterpret_cast<const double2*>(bsrc + 16 * j);
                c23 = *reinterpret_cast<const double2*>(bsrc + 16 * j + 2);
            }
            dmma884(a0, a1, (double)fa.x, c01.x); dmma884(b0, b1, (double)fb.x, c01.x);
            dmma884(a0, a1, (double)fa.y, c01.y); dmma884(b0, b1, (double)fb.y, c01.y);
            dmma884(a0, a1, (double)fa.z, c23.x); dmma884(b0, b1, (double)fb.z, c23.x);
            dmma884(a0, a1, (double)fa.w, c23.y); dmma884(b0, b1, (double)fb.w, c23.y);
        }
        // lane holds x_row . c_t for t = 2q, 2q + 1 (rows ra and rb)
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int64_t row = g ? rb : ra;
            if (row < N && t0 < T) {
                const double xs = __ldg(xsq + row);
                const float prev = closest ? __ldg(closest + row) : INFINITY;
                // ((-2 x.c) + |c|^2) + |x|^2 in fp64, then fp32, then clamp (pairwise.py:632-636, 413)
                float dist = (float)((-2.0 * (g ? b0 : a0) + s_cc[t0]) + xs);
                dist = fminf(fmaxf(dist, 0.f), prev);
                out[(size_t)t0 * N + row] = dist;
                pot0 += (double)dist;
                if (t1 < T) {
                    float dist1 = (float)((-2.0 * (g ? b1 : a1) + s_cc[t1]) + xs);
                    dist1 = fminf(fmaxf(dist1, 0.f), prev);
                    out[(size_t)t1 * N + row] = dist1;
                    pot1 += (double)dist1;
                }
            }
        }
    }
    // lanes with the same q hold partial potentials of the same two candidates
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
        pot0 += __shfl_xor_sync(0xffffffffu, pot0, o);
        pot1 += __shfl_xor_sync(0xffffffffu, pot1, o);
    }
    if (lane < 4) { s_pot[warp][2 * lane] = pot0; s_pot[warp][2 * lane + 1] = pot1; }
    __syncthreads();
    if (threadIdx.x < T) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += s_pot[w][threadIdx.x];
        part[(size_t)blockIdx.x * T + threadIdx.x] = s;
    }
}

static int kpp_grid() { return num_sms() * 4; }

}  // namespace vtf

using namespace vtf;


// ---- round glue: everything SK/cluster/_kmeans.py:249-266 does between two distance passes, on the device, so that a
// k-means++ round needs no host round trip.  One thread; T <= kMaxTrials.
//   in : pots[n_pots]  fp64 potentials of the previous round's candidates (n_pots = 1 after the first centre)
//        u[T]          the round's uniform draws (host MT19937, drawn in advance: they do not depend on the data)
//   out: best          argmin of the fp32-rounded potentials (first minimum, np.argmin), :266
//        thr[T]        ascending fp32 thresholds = smallest fp32 >= u * pot (searchsorted compares the fp32 cumsum with
//                      fp64 values, :249-250), order[T] = draw index of each sorted threshold
__global__ void kpp_round_prep_kernel(const double* __restrict__ pots, int n_pots, const double* __restrict__ u, int T,
                                      float* __restrict__ thr, int* __restrict__ order, int64_t* __restrict__ best) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int b = 0;
    float pot = (float)pots[0];
    for (int i = 1; i < n_pots; ++i) {
        const float p = (float)pots[i];
        if (p < pot) { pot = p; b = i; }
    }
    if (best) *best = b;
    for (int t = 0; t < T; ++t) {
        const double v = u[t] * (double)pot;
        float f = __double2float_rn(v);
        if ((double)f < v) f = nextafterf(f, INFINITY);
        int j = t;                                               // stable insertion sort
        while (j > 0 && thr[j - 1] > f) { thr[j] = thr[j - 1]; order[j] = order[j - 1]; --j; }
        thr[j] = f;
        order[j] = t;
    }
}

// cand[order[t]] = picked[t], never-reached thresholds clip to N-1 (_kmeans.py:254)
__global__ void kpp_unsort_kernel(const int64_t* __restrict__ picked, const int* __restrict__ order, int T, int64_t N,
                                  int64_t* __restrict__ cand) {
    const int t = threadIdx.x;
    if (t < T) cand[order[t]] = picked[t] < 0 ? N - 1 : picked[t];
}

extern "C" {

size_t vtf_kpp_ws_bytes(int T) { return (size_t)kpp_grid() * (size_t)(T > 0 ? T : 1) * sizeof(double); }

int vtf_kpp_rownorm64(const float* Xc, int64_t N, int d, double* xsq, void* stream) {
    VTF_REQUIRE(Xc && xsq && N >= 0 && d > 0 && d % 4 == 0, "vtf_kpp_rownorm64");
    if (N == 0) return VTF_OK;
    const int grid = (int)min((int64_t)num_sms() * 8, (N + 7) / 8);
    VTF_LAUNCH(kpp_rownorm64_kernel, grid, 256, 0, stream, Xc, N, d, xsq);
    VTF_CHECK_LAST("vtf_kpp_rownorm64");
    return VTF_OK;
}

int vtf_kpp_dist(const float* Xc, int64_t N, int d, const float* cand, int T, const float* closest, const double* xsq, float* out,
                 double* pots, void* ws, void* stream) {
    VTF_REQUIRE(Xc && cand && out && pots && ws && N >= 0 && d > 0 && d % 4 == 0 && T > 0 && T <= kMaxTrials,
                "vtf_kpp_dist");
    const size_t smem = (size_t)T * d * sizeof(double);
    VTF_REQUIRE(smem <= 128 * 1024, "vtf_kpp_dist: T*d too large");
    const int grid = kpp_grid();
#define VTF_KPP_LAUNCH(TTv, Rv, Hv)                                                                                          \
    do {                                                                                                                      \
        VTF_CUDA(cudaFuncSetAttribute(kpp_dist_kernel<TTv, Rv, Hv>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024), \
                 "vtf_kpp_dist");                                                                                             \
        VTF_LAUNCH((kpp_dist_kernel<TTv, Rv, Hv>), grid, 256, smem, stream, Xc, N, d, cand, T, closest, xsq, out, (double*)ws); \
    } while (0)
    static int use_mma = -1;
    if (use_mma < 0) { const char* e = getenv("VTF_KPP_MMA"); use_mma = e ? atoi(e) : 1; }
    const size_t smem_mma = (size_t)T * (d + 2) * sizeof(double);
    if (use_mma && xsq && d % 16 == 0 && smem_mma <= 100 * 1024) {
        VTF_CUDA(cudaFuncSetAttribute(kpp_dist_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024), "vtf_kpp_dist");
        // latency-bound on the X loads (ncu: long_scoreboard): three resident CTAs per SM, and a grid that fills them evenly
        static int mma_grid = 0;
        if (!mma_grid) { const char* e = getenv("VTF_KPP_MMA_GRID"); mma_grid = e ? atoi(e) : 3; }
        const int g = min(grid, num_sms() * mma_grid);
        VTF_LAUNCH(kpp_dist_mma_kernel, g, 256, smem_mma, stream, Xc, N, d, cand, T, closest, xsq, out, (double*)ws);
        VTF_CHECK_LAST("vtf_kpp_dist");
        VTF_LAUNCH(reduce_parts_kernel, 1, 32, 0, stream, (const double*)ws, g, T, pots, pots, T);
        VTF_CHECK_LAST("vtf_kpp_dist/reduce");
        return VTF_OK;
    } else if (T <= 4) { if (xsq) VTF_KPP_LAUNCH(4, 4, true); else VTF_KPP_LAUNCH(4, 4, false); }
    else { if (xsq) VTF_KPP_LAUNCH(kMaxTrials, 2, true); else VTF_KPP_LAUNCH(kMaxTrials, 2, false); }
#undef VTF_KPP_LAUNCH
    VTF_CHECK_LAST("vtf_kpp_dist");
    VTF_LAUNCH(reduce_parts_kernel, 1, 32, 0, stream, (const double*)ws, grid, T, pots, pots, T);
    VTF_CHECK_LAST("vtf_kpp_dist/reduce");
    return VTF_OK;
}

int vtf_kpp_pick_serial(const float* closest, int64_t N, const float* thr, int T, const float* offset_in, float* total_out,
                        int64_t* idx, void* stream) {
    VTF_REQUIRE(closest && thr && idx && N >= 0 && T > 0 && T <= kMaxTrials, "vtf_kpp_pick_serial");
    VTF_LAUNCH(kpp_pick_kernel, 1, 128, 0, stream, closest, N, thr, T, offset_in, total_out, idx);
    VTF_CHECK_LAST("vtf_kpp_pick_serial");
    return VTF_OK;
}

int vtf_kpp_pick(const float* closest, int64_t N, const float* thr, int T, const float* offset_in, float* total_out,
                 int64_t* idx, void* stream) {
    VTF_REQUIRE(closest && thr && idx && N >= 0 && T > 0 && T <= kMaxTrials, "vtf_kpp_pick");
    static int serial = -1;
    if (serial < 0) { const char* e = getenv("VTF_PICK_SERIAL"); serial = (e && atoi(e)) ? 1 : 0; }
    if (serial) VTF_LAUNCH(kpp_pick_kernel, 1, 128, 0, stream, closest, N, thr, T, offset_in, total_out, idx);
    else VTF_LAUNCH(kpp_pick_scan_kernel<false>, 1, kScanThreads, 0, stream, closest, N, thr, T, offset_in, total_out, idx,
                    (const long long*)nullptr, (const int*)nullptr);
    VTF_CHECK_LAST("vtf_kpp_pick");
    return VTF_OK;
}

size_t vtf_kpp_pick_ws_bytes(int64_t N) {
    const size_t tiles = (size_t)((N + kScanTile - 1) / kScanTile) + 1;
    return tiles * (sizeof(double) + sizeof(long long) + sizeof(int)) + 64;
}

int vtf_kpp_pick_fast(const float* closest, int64_t N, const float* thr, int T, const float* offset_in, float* total_out,
                      int64_t* idx, void* ws, void* stream) {
    VTF_REQUIRE(closest && thr && idx && ws && N >= 0 && T > 0 && T <= kMaxTrials, "vtf_kpp_pick_fast");
    const int64_t tiles = (N + kScanTile - 1) / kScanTile;
    if (tiles < 4) {                                 // nothing to skip
        VTF_LAUNCH(kpp_pick_scan_kernel<false>, 1, kScanThreads, 0, stream, closest, N, thr, T, offset_in, total_out, idx,
                   (const long long*)nullptr, (const int*)nullptr);
        VTF_CHECK_LAST("vtf_kpp_pick_fast");
        return VTF_OK;
    }
    double* csum = (double*)ws;
    long long* cc = (long long*)(csum + tiles + 1);
    int* ce = (int*)(cc + tiles + 1);
    VTF_LAUNCH(kpp_chunk_sum_kernel, (unsigned)tiles, 256, 0, stream, closest, N, csum);
    VTF_CHECK_LAST("vtf_kpp_pick_fast/sum");
    VTF_LAUNCH(kpp_chunk_int_kernel, (unsigned)tiles, 256, 0, stream, closest, N, (const double*)csum, offset_in, cc, ce);
    VTF_CHECK_LAST("vtf_kpp_pick_fast/int");
    VTF_LAUNCH(kpp_pick_scan_kernel<true>, 1, kScanThreads, 0, stream, closest, N, thr, T, offset_in, total_out, idx,
               (const long long*)cc, (const int*)ce);
    VTF_CHECK_LAST("vtf_kpp_pick_fast");
    return VTF_OK;
}

int vtf_kpp_round_prep(const double* pots, int n_pots, const double* u, int T, float* thr, int32_t* order, int64_t* best,
                       void* stream) {
    VTF_REQUIRE(pots && u && thr && order && n_pots > 0 && T > 0 && T <= kMaxTrials, "vtf_kpp_round_prep");
    VTF_LAUNCH(kpp_round_prep_kernel, 1, 32, 0, stream, pots, n_pots, u, T, thr, order, best);
    VTF_CHECK_LAST("vtf_kpp_round_prep");
    return VTF_OK;
}

int vtf_kpp_unsort(const int64_t* picked, const int32_t* order, int T, int64_t N, int64_t* cand, void* stream) {
    VTF_REQUIRE(picked && order && cand && T > 0 && T <= kMaxTrials && N > 0, "vtf_kpp_unsort");
    VTF_LAUNCH(kpp_unsort_kernel, 1, 32, 0, stream, picked, order, T, N, cand);
    VTF_CHECK_LAST("vtf_kpp_unsort");
    return VTF_OK;
}

}  // extern "C"
