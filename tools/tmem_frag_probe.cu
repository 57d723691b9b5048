// Probe (not part of the library; round-2 groundwork, NOT yet run on hardware): which TMEM
// (lane, column) does each register of a tcgen05.ld.16x256b fragment hold?  The accumulation epilogue
// (tc_accum.cu) wants one thread to own the two reduced one-hot planes of a column; if the 16x256b
// fragment gives thread t lanes t/4 and t/4 + 8 (the mma-style C fragment), planes placed 8 one-hot rows
// apart meet in one thread.
//   make -C fastselect_b200/csrc probe3 && tools/build/tmem_frag_probe
// TMEM is filled with lane * 1000 + column through the 32x32b shape (thread = lane), read back with
// 16x256b.x1 by warp 0, and the mapping is printed.
#include <cstdio>
#include <cuda_runtime.h>
#include "../fastselect_b200/csrc/tc_common.cuh"

using namespace fs::tc;

__global__ void __launch_bounds__(128, 1) probe(uint32_t *out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<32>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot;
    for (int c = 0; c < 32; ++c)      // every warp fills its own lane quarter: value = lane * 1000 + column
        tmem_st_32x1(base + ((uint32_t)(warp * 32) << 16) + c, (uint32_t)((warp * 32 + lane) * 1000 + c));
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        uint32_t r[4];
        // lanes 0..15 of this warp's quarter, columns 0..7 (256 bits per lane)
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(base)
                     : "memory");
        tmem_ld_wait();
        for (int i = 0; i < 4; ++i) out[lane * 4 + i] = r[i];
        // and lanes 16..31 (lane offset 16 in the address)
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(base + (16u << 16))
                     : "memory");
        tmem_ld_wait();
        for (int i = 0; i < 4; ++i) out[128 + lane * 4 + i] = r[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<32>(base);
    }
}

int main() {
    uint32_t *d, h[256];
    if (cudaMalloc(&d, sizeof(h)) != cudaSuccess) return 1;
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        fprintf(stderr, "probe failed: %s\n", cudaGetErrorString(e));
        return 1;
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int half = 0; half < 2; ++half) {
        printf("16x256b.x1 at lane offset %d: thread -> (lane, column) of r0..r3\n", 16 * half);
        for (int t = 0; t < 32; ++t) {
            printf("  t%02d:", t);
            for (int i = 0; i < 4; ++i) printf("  (%3u,%2u)", h[half * 128 + t * 4 + i] / 1000, h[half * 128 + t * 4 + i] % 1000);
            printf("\n");
        }
    }
    return 0;
}
