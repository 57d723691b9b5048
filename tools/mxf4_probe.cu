// Probe (not part of the library): is tcgen05.mma kind::mxf4 (e2m1 operands, UE8M0 block scales
// of 1.0, FP32 accumulation) EXACT for small-integer operands up to sums of 2^22, and how fast is
// it next to kind::i8?  Operand tiles are uniform, so their shared-memory layout does not matter.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/mxf4_probe tools/mxf4_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../fastselect_b200/csrc/tc_common.cuh"

using namespace fs::tc;

__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) { tmem_st_32x1(taddr, v); }

constexpr int BM = 128, BN = 256;
// mode 0: mxf4, mode 1: i8
__global__ void __launch_bounds__(128, 1) probe(int mode, int iters, uint32_t a_byte, uint32_t b_byte, float *out, long long *cycles, int nsf, int sfb_off, int bn) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sa = smem, *sb = smem + BM * 128;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + BM * 128 + BN * 128);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < BM * 128 / 4; i += 128) reinterpret_cast<uint32_t *>(sa)[i] = a_byte * 0x01010101u;
    for (int i = threadIdx.x; i < BN * 128 / 4; i += 128) reinterpret_cast<uint32_t *>(sb)[i] = b_byte * 0x01010101u;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *slot;
    // scale factors: every byte 0x7f (UE8M0 1.0) in columns 256..287 of all lanes
    // only the first nsf columns hold valid scales; the rest are 2^-127 so that any read beyond shows up
    for (int c = 0; c < 32; ++c) tmem_st1(tbase + ((uint32_t)(warp * 32) << 16) + 256 + c, c < nsf ? 0x7f7f7f7fu : 0u);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        const uint64_t da = make_smem_desc_sw128(smem_u32(sa)), db = make_smem_desc_sw128(smem_u32(sb));
        // block-scaled descriptor: a/b format e2m1 (1), UE8M0 scales, N >> 3, M >> 4
        const uint32_t idesc4 = (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | (1u << 23) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t idesc8 = make_idesc_i8(BM, BN);
        t0 = clock64();
        for (int it = 0; it < iters; ++it)
            for (int k = 0; k < 4; ++k) {
                if (mode == 0) mma_mxf4(tbase, da + 2 * k, db + 2 * k, idesc4, tbase + 256, tbase + 256 + sfb_off, (it | k) != 0);
                else mma_i8(tbase, da + 2 * k, db + 2 * k, idesc8, (it | k) != 0);
            }
        tc_commit(bar);
    }
    mbar_wait(bar, 0);
    if (threadIdx.x == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
    tc_fence_after();
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int e = 0; e < 32; ++e) {
            const float f = mode == 0 ? __uint_as_float(v[e]) : (float)(int)v[e];
            out[(size_t)(warp * 32 + lane) * BN + c0 + e] = f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tbase); }
}

int main() {
    const int smem = BM * 128 + BN * 128 + 1024 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float *out; long long *cyc;
    cudaMalloc(&out, BM * BN * sizeof(float)); cudaMalloc(&cyc, 8);
    std::vector<float> h(BM * BN);
    // how many TMEM columns do the (all-ones) scale factors of A (M=128) and B (N=bn) occupy?
    for (int bn : {256, 240, 192})
        for (int sfb_off : {0, 4})
            for (int nsf : {1, 2, 3, 4, 6, 8}) {
                cudaMemset(out, 0xff, BM * BN * sizeof(float));
                probe<<<1, 128, smem>>>(0, 64, 0x22, 0x44, out, cyc, nsf, sfb_off, bn);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
                long bad = 0;
                for (int r = 0; r < BM; ++r)
                    for (int c = 0; c < bn; ++c) bad += (double)h[(size_t)r * BN + c] != 2.0 * 256 * 64;
                long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
                printf("N=%d sfb_off=%d valid scale columns=%d: mismatches %ld  cycles/MMA %.1f\n", bn, sfb_off, nsf, bad, (double)hc / 256.0);
            }
    return 0;
}
