// Probe: which (TMEM lane, column) lands in which (thread, register) for tcgen05.ld.16x256b.x4?
// Warp 0 writes value = lane * 1000 + column with tcgen05.st.32x32b.x32 (thread i = lane i, register j = column j),
// reads the same 32 x 32 block back as two 16x256b.x4 loads (lanes 0-15 and 16-31) and prints what every thread got.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe/tmem_ld_layout tools/probe/tmem_ld_layout.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(int* out) {
    __shared__ uint32_t slot;
    const int lane = threadIdx.x;
    const uint32_t slot_addr = (uint32_t)__cvta_generic_to_shared(&slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_addr), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = lane * 1000 + j;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(base), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int half = 0; half < 2; ++half) {
        uint32_t r[16];
        const uint32_t addr = base + ((uint32_t)(half * 16) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(addr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[(half * 32 + lane) * 16 + j] = (int)r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32u) : "memory");
}

int main() {
    int* d;
    cudaMalloc(&d, 2 * 32 * 16 * sizeof(int));
    probe<<<1, 32>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    static int h[2 * 32 * 16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int half = 0; half < 2; ++half)
        for (int t = 0; t < 32; ++t) {
            printf("half %d thread %2d:", half, t);
            for (int j = 0; j < 16; ++j) {
                const int v = h[(half * 32 + t) * 16 + j];
                printf(" (%d,%d)", v / 1000, v % 1000);
                // expectation = the mma.m16n8 accumulator layout repeated over 8-column groups
                const int row = half * 16 + t / 4 + ((j & 2) ? 8 : 0), col = (j >> 2) * 8 + (t % 4) * 2 + (j & 1);
                if (v != row * 1000 + col) ++bad;
            }
            printf("\n");
        }
    printf("mismatches against the assumed layout (row = 16*half + t/4 + 8*((j>>1)&1), col = 8*(j>>2) + 2*(t%%4) + (j&1)): %d\n", bad);
    return 0;
}
