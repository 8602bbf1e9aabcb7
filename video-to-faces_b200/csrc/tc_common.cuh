// tcgen05 / TMA / mbarrier building blocks shared by the tensor-core kernels (sm_100a inline PTX).
// PTX forms follow cute/arch/{mma_sm100_umma,copy_sm100_tma,tmem_allocator_sm100}.hpp and cutlass/arch/barrier.h.
#pragma once
#include <cuda.h>
#include "common.cuh"

#ifndef VTF_CELL_PREFETCH
#define VTF_CELL_PREFETCH 1
#endif
#ifndef VTF_WAIT_HINT_NS
#define VTF_WAIT_HINT_NS 2000   // suspend-time hint of mbarrier.try_wait: fewer polls, less power, higher clock (-2.5%)
#endif
#ifndef VTF_TIMEOUT_EVERY
#define VTF_TIMEOUT_EVERY 1      // read clock64() for the hang watchdog every n-th failed poll
#endif

namespace vtf {
namespace tc {

constexpr int BM = 128;          // rows per CTA tile (TMEM lanes)
constexpr int BN = 256;          // columns per MMA tile (TMEM columns per accumulator stage)
constexpr int BK = 64;           // fp16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 320;    // 2 control warps + 8 epilogue warps
constexpr int kABlockBytes = BM * BK * 2;   // 16 KB
constexpr int kBBlockBytes = BN * BK * 2;   // 32 KB
constexpr unsigned long long kTimeoutCycles = 6000000000ull;   // ~3 s: trap instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, polls = 0;
    const unsigned long long t0 = clock64();
    while (true) {
#if VTF_WAIT_HINT_NS > 0
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity), "r"((uint32_t)VTF_WAIT_HINT_NS) : "memory");
#else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
#endif
        if (done) return;
        if ((++polls % VTF_TIMEOUT_EVERY) == 0 && clock64() - t0 > kTimeoutCycles) __trap();   // a broken pipeline must not hang the device
    }
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// ---- CTA-pair (cta_group::2) flavours: the pair shares one 256-row MMA, each CTA feeds half of B
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same offset in CTA `cta` of the cluster.  RELAXED drops .release.cluster (the
// cutlass ClusterBarrier::arrive form): the qualifier costs a membar (ERRBAR) per arrive, and the only thing ordered
// here is TMEM reads, which tcgen05.wait::ld + tcgen05.fence::before_thread_sync already cover.  Which form is faster
// is a power question and was settled by same-box A/B runs (tools/sweep_ab.py): the full sweep is 5% faster WITH
// the membar (825 vs 971 ms before the wait hint existed), the symmetric sweep 10% faster WITHOUT it (407 vs 452 ms).
template <bool RELAXED>
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    if (RELAXED) {
        asm volatile(
            "{\n\t.reg .b32 ra;\n\t"
            "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
            "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
            ::"r"(bar), "r"(cta) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .b32 ra;\n\t"
            "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
            "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
            ::"r"(bar), "r"(cta) : "memory");
    }
}
// wait with acquire at CLUSTER scope: the data guarded by the barrier was written by the other CTA of the pair
// through distributed shared memory (st.shared::cluster + mbarrier.arrive.release.cluster)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const unsigned long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity), "r"(2000u) : "memory");
        if (done) return;
        if (clock64() - t0 > kTimeoutCycles) __trap();
    }
}
// store a word at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void st_shared_remote_u32(uint32_t addr, uint32_t cta, uint32_t v) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.u32 [ra], %2;\n\t}"
        ::"r"(addr), "r"(cta), "r"(v) : "memory");
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // cute::Sm100MmaPeerBitMask: address of the same barrier in the even CTA
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T, fp16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// K-major operand, 128-byte swizzle, rows 128 B apart, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor:
// start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64))
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// The augmented 16-column block is stored in the canonical NO-swizzle K-major order (silhouette.cu, store_aug):
// core matrix = 8 rows x 16 B contiguous, the two K core matrices 128 B apart (LBO), 8-row groups 256 B apart (SBO).
__device__ __forceinline__ uint64_t make_desc_aug(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}

// cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format F16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t kIdescPair = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);   // M = 256 over two CTAs

#define VTF_TMEM_LD32(taddr, v)                                                                              \
    asm volatile(                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"    \
        "%26,%27,%28,%29,%30,%31}, [%32];"                                                                   \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),    \
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),           \
          "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),         \
          "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),         \
          "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                              \
        : "r"(taddr))

// 16 lanes x 256 bits, 4 repetitions = 16 lanes x 32 columns in the mma.m16n8 accumulator layout: thread (g = lane / 4,
// tq = lane % 4) receives, for column group i = 0..3, v[4i + {0,1}] = row g, columns 8i + 2tq + {0,1} and
// v[4i + {2,3}] = row g + 8, same columns (verified on the device by tools/probe/tmem_ld_layout.cu).
#define VTF_TMEM_LD16X256(taddr, v, o)                                                                       \
    asm volatile(                                                                                            \
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "                                                            \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                                    \
        : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),     \
          "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),  \
          "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])                                  \
        : "r"(taddr))

// the "+r" operands tie every later use of v[] to the completion of the load
#define VTF_TMEM_WAIT(v)                                                                                     \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                            \
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),       \
                   "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]),   \
                   "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]),             \
                   "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]),             \
                   "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]))

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// [P, d] fp16 row-major operand, boxes of 64 columns x box_rows rows, 128-byte swizzle
inline int make_operand_map(CUtensorMap* map, const void* Xh, int64_t P, int d, int box_rows, const char* who) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(VTF_ERR_CUDA, who, "cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)P};
    cuuint64_t strides[1] = {(cuuint64_t)d * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(Xh), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[64];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d)", (int)r);
        return fail(VTF_ERR_CUDA, who, buf);
    }
    return VTF_OK;
}

}  // namespace tc
}  // namespace vtf
