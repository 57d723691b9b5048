// Whole-GPU peak of tcgen05.mma kind::mxf4 (e2m1 operands, unit UE8M0 block scales, FP32 accumulation
// in TMEM) -- the denominator of the tensor-pipe roofline in bench.py.  MEASURED_PEAKS.json has no FP4
// figure, so this tool measures one on the same GPU: one CTA (or CTA pair) per SM, operand tiles
// resident in shared memory (no TMA, no global traffic, no epilogue), one thread issuing back-to-back
// MMAs into a TMEM accumulator.  This is the rate the tensor pipe sustains when nothing but the
// shared-memory operand reads limits it; a real GEMM cannot exceed it.
//
//   make -C fastselect_b200/csrc peak && tools/build/fp4_peak [seconds_sustained]
//
// Reports, per shape (M128 x N per CTA, cta_group::1; M256 x N per CTA pair, cta_group::2):
//   cycles per MMA on one SM (clock64), whole-GPU POP/s as a burst (one ~2 ms launch, best of 5) and
//   sustained (back-to-back launches for `seconds_sustained`, default 4 s, rate of the second half).
// Output is plain text plus one JSON line (profiles/r02_fp4_peak.json is a copy of it).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../fastselect_b200/csrc/tc_common.cuh"

using namespace fs::tc;

namespace {
constexpr int BM = 128;
constexpr int SF_COL = 256;

__device__ __forceinline__ void cluster_sync() { cluster_sync_all(); }
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// bounded wait (5 s): a protocol bug traps instead of hanging the GPU box
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    const uint64_t t0 = global_ns();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if (global_ns() - t0 > 5000000000ull) break;
    }
    printf("fp4_peak: mbarrier timeout (block %d)\n", (int)blockIdx.x);
    __trap();
}
__device__ __forceinline__ void tmem_alloc_pair512(uint32_t *dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(dst_smem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair512(uint32_t addr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}
}  // namespace

// cta_group::1: every CTA owns a 128 x bn accumulator; A tile 128 rows x 128 B, B tile bn rows x 128 B.
__global__ void __launch_bounds__(128, 1) peak_cg1(int iters, int bn, long long *cycles, float *check) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sa = smem, *sb = smem + BM * 128;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + BM * 128 + 256 * 128);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < BM * 128 / 4; i += 128) reinterpret_cast<uint32_t *>(sa)[i] = 0x22222222u;   // 1.0
    for (int i = threadIdx.x; i < 256 * 128 / 4; i += 128) reinterpret_cast<uint32_t *>(sb)[i] = 0x22222222u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *slot;
    for (int c = 0; c < 8; ++c) tmem_st_32x1(tbase + ((uint32_t)(warp * 32) << 16) + SF_COL + c, 0x7f7f7f7fu);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint64_t da = make_smem_desc_sw128(smem_u32(sa)), db = make_smem_desc_sw128(smem_u32(sb));
        const uint32_t idesc = make_idesc_mxf4(BM, bn);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_mxf4(tbase, da + 2 * k, db + 2 * k, idesc, tbase + SF_COL, tbase + SF_COL, (it | k) != 0);
        tc_commit(bar);
        mbar_wait_bounded(bar, 0);
        if (blockIdx.x == 0) cycles[0] = clock64() - t0;
    }
    __syncthreads();
    tc_fence_after();
    if (blockIdx.x == 0 && check) {
        uint32_t v[16];
        tmem_ld_32x16(tbase + ((uint32_t)(warp * 32) << 16), v);
        tmem_ld_wait();
        if (threadIdx.x == 0) check[0] = __uint_as_float(v[0]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<512>(tbase);
    }
}

// cta_group::2: a CTA pair owns a 256 x bn accumulator (128 rows per CTA); every CTA holds its own
// 128 A rows and bn / 2 B rows; the leader issues for the pair.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) peak_cg2(int iters, int bn, long long *cycles, float *check) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sa = smem, *sb = smem + BM * 128;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + BM * 128 + 256 * 128);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < BM * 128 / 4; i += 128) reinterpret_cast<uint32_t *>(sa)[i] = 0x22222222u;
    for (int i = threadIdx.x; i < 256 * 128 / 4; i += 128) reinterpret_cast<uint32_t *>(sb)[i] = 0x22222222u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc_pair512(slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const uint32_t tbase = *slot;
    for (int c = 0; c < 8; ++c) tmem_st_32x1(tbase + ((uint32_t)(warp * 32) << 16) + SF_COL + c, 0x7f7f7f7fu);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    if (threadIdx.x == 0 && rank == 0) {
        const uint64_t da = make_smem_desc_sw128(smem_u32(sa)), db = make_smem_desc_sw128(smem_u32(sb));
        const uint32_t idesc = make_idesc_mxf4(2 * BM, bn);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                mma_mxf4_pair(tbase, da + 2 * k, db + 2 * k, idesc, tbase + SF_COL, tbase + SF_COL, (it | k) != 0);
        tc_commit_pair(bar);
        mbar_wait_bounded(bar, 0);
        if (blockIdx.x == 0) cycles[0] = clock64() - t0;
    } else if (threadIdx.x == 0) {
        mbar_wait_bounded(bar, 0);       // the multicast commit arrives here too
    }
    __syncthreads();
    tc_fence_after();
    if (blockIdx.x == 0 && check) {
        uint32_t v[16];
        tmem_ld_32x16(tbase + ((uint32_t)(warp * 32) << 16), v);
        tmem_ld_wait();
        if (threadIdx.x == 0) check[0] = __uint_as_float(v[0]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc_pair512(tbase);
    }
}

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s failed at line %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

struct Result {
    int cg, bn;
    double cycles_per_mma, burst_pops, sustained_pops;
    float check;
};

static Result measure(int cg, int bn, int sms, double seconds) {
    const int smem = BM * 128 + 256 * 128 + 1024 + 64;
    long long *cyc;
    float *chk;
    CK(cudaMalloc(&cyc, 8));
    CK(cudaMalloc(&chk, 4));
    CK(cudaMemset(cyc, 0, 8));
    if (cg == 1) CK(cudaFuncSetAttribute(peak_cg1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else CK(cudaFuncSetAttribute(peak_cg2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = cg == 1 ? sms : (sms / 2) * 2;
    auto launch = [&](int iters) {
        if (cg == 1) peak_cg1<<<grid, 128, smem>>>(iters, bn, cyc, chk);
        else peak_cg2<<<grid, 128, smem>>>(iters, bn, cyc, chk);
        CK(cudaGetLastError());
    };
    // ops of one launch: every CTA (cg1) or CTA pair (cg2) issues 4 * iters MMAs of M x bn x 64
    auto ops = [&](int iters) {
        const double per_mma = 2.0 * (cg == 1 ? BM : 2 * BM) * bn * 64.0;
        return per_mma * 4.0 * iters * (cg == 1 ? grid : grid / 2);
    };
    const int iters = 6000;              // ~2 ms per launch
    launch(200);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        launch(iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    long long hc = 0;
    float hk = 0.f;
    CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&hk, chk, 4, cudaMemcpyDeviceToHost));
    Result res{cg, bn, (double)hc / (4.0 * iters), ops(iters) / (best * 1e-3) / 1e15, 0.0, hk};
    if (seconds > 0) {
        // sustained: launches back to back; the rate of the second half of the interval
        const int n_launch = (int)(seconds * 1e3 / best) + 2;
        const int half = n_launch / 2;
        cudaEvent_t em;
        CK(cudaEventCreate(&em));
        for (int r = 0; r < n_launch; ++r) {
            if (r == half) CK(cudaEventRecord(em));
            launch(iters);
        }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, em, e1));
        res.sustained_pops = ops(iters) * (n_launch - half) / (ms * 1e-3) / 1e15;
    }
    cudaFree(cyc);
    cudaFree(chk);
    return res;
}

int main(int argc, char **argv) {
    const double seconds = argc > 1 ? atof(argv[1]) : 4.0;
    int sms = 0, clk_khz = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("# fp4_peak: %d SMs, max SM clock %.0f MHz; tcgen05.mma kind::mxf4.block_scale, operands resident in shared memory\n", sms,
           clk_khz / 1e3);
    std::vector<Result> all;
    const int shapes[][2] = {{1, 256}, {1, 240}, {1, 192}, {1, 128}, {2, 256}, {2, 240}, {2, 128}};
    for (auto &s : shapes) {
        // the sustained leg only for the two shapes the library's kernels use as their denominators
        const bool sustain = s[1] == 256;
        Result r = measure(s[0], s[1], sms, sustain ? seconds : 0.0);
        all.push_back(r);
        printf("cta_group::%d M%d N%d K64: %.1f cycles/MMA on one SM, burst %.3f POP/s", r.cg, r.cg * BM, r.bn, r.cycles_per_mma,
               r.burst_pops);
        if (sustain) printf(", sustained (%.0f s) %.3f POP/s", seconds, r.sustained_pops);
        printf("  [accumulator check: %.0f, expected %.0f]\n", r.check, 6000.0 * 4 * 64);
    }
    printf("{\"tool\": \"fp4_peak\", \"sms\": %d, \"sm_max_mhz\": %.0f, \"shapes\": [", sms, clk_khz / 1e3);
    for (size_t i = 0; i < all.size(); ++i)
        printf("%s{\"cta_group\": %d, \"m\": %d, \"n\": %d, \"cycles_per_mma\": %.2f, \"burst_pops\": %.4f, \"sustained_pops\": %.4f}",
               i ? ", " : "", all[i].cg, all[i].cg * BM, all[i].bn, all[i].cycles_per_mma, all[i].burst_pops, all[i].sustained_pops);
    printf("]}\n");
    return 0;
}
