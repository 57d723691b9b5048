// Probe (not part of the library; round-2 groundwork, NOT yet run on hardware): the distance GEMM
// D = A * B^T on FP4 (e2m1) operands as a CTA PAIR -- tcgen05.mma.cta_group::2, M = 256 (128 rows per
// CTA) x N = 256, each CTA staging its own 128 A rows and HALF of the 256 B rows, so a pair moves
// 32 KB of operands per CTA and K slab instead of the 48 KB of the single-CTA kernel in
// fastselect_b200/csrc/tc_dist.cu (whose MMA issuer waits for operands: DESIGN.md section 8).
//
//   make -C fastselect_b200/csrc probe2 && tools/build/tc_dist_cg2_probe [M N Kbytes]
//
// Checks the int32 result against a CPU GEMM at a small shape (this also settles which B rows /
// D columns belong to which CTA of the pair) and times a large shape.  Every mbarrier wait is
// bounded: a protocol bug traps instead of hanging the GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../fastselect_b200/csrc/tc_common.cuh"

// (stand-alone probe from before these helpers moved into tc_common.cuh: it keeps its own copies under other names)
using namespace fs::tc;

namespace {
constexpr int BM = 128;            // A rows per CTA (UMMA M = 256 over the pair)
constexpr int BN = 256;            // D columns per pair
constexpr int BNH = BN / 2;        // B rows staged by each CTA
constexpr int BK = 128;            // bytes of K per stage (one 128-byte swizzle atom = 256 nibbles)
constexpr int STAGES = 6;
constexpr int A_BYTES = BM * BK;   // 16 KB
constexpr int B_BYTES = BNH * BK;  // 16 KB
constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 + 256;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int SF_COL = BN;         // unit UE8M0 scales in columns 256..287
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address: the pair's leader

__device__ __forceinline__ uint32_t probe_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void probe_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bounded wait: two seconds of wall time, then trap (a hang would cost the whole GPU box)
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity, int tag) {
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
        if (global_ns() - t0 > 2000000000ull) break;
    }
    printf("mbarrier timeout: tag %d block %d thread %d parity %u\n", tag, (int)blockIdx.x, (int)threadIdx.x, parity);
    __trap();
}
// 2-D tile load issued by either CTA of the pair; completion bytes go to the LEADER's mbarrier
__device__ __forceinline__ void probe_tma_load_2d_pair(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void probe_tmem_alloc_pair(uint32_t *dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void probe_tmem_dealloc_pair(uint32_t addr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(TMEM_COLS) : "memory");
}
// arrives on the mbarrier at this shared-memory offset in BOTH CTAs when the pair's MMAs retire
__device__ __forceinline__ void probe_tc_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void probe_mma_mxf4_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}
}  // namespace

// One CTA pair per 256 x 256 tile of D.  blockIdx.x = 2 * tile + rank.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
dist_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int num_k_blocks,
                 int32_t *__restrict__ D, int64_t M, int64_t N, int64_t ldd, int tiles_x) {
    const uint32_t rank = probe_cluster_ctarank();
    const int tile = (int)(blockIdx.x >> 1);
    const int tile_y = tile / tiles_x, tile_x = tile % tiles_x;
    const int64_t m0 = (int64_t)tile_y * 2 * BM + rank * BM;      // this CTA's A rows = its D rows
    const int64_t n0 = (int64_t)tile_x * BN;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *smem_a = smem;
    unsigned char *smem_b = smem + STAGES * A_BYTES;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * (A_BYTES + B_BYTES));
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *accum_bar = empty_bar + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);      // the leader's producer arrives (with the pair's byte count)
            mbar_init(&empty_bar[s], 1);     // the pair's MMA commit arrives (multicast)
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) probe_tmem_alloc_pair(tmem_slot);     // same warp id in both CTAs, same slot offset
    tc_fence_before();
    __syncthreads();
    probe_cluster_sync();                                // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 2) {
        for (int c = 0; c < 32; ++c) tmem_st_32x1(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + SF_COL + c, 0x7f7f7f7fu);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    probe_cluster_sync();                                // both CTAs' scale factors are in place before the leader issues
    tc_fence_after();

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait_bounded(&empty_bar[s], ph ^ 1, 100 + s);
                if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (A_BYTES + B_BYTES));
                probe_tma_load_2d_pair(smem_a + s * A_BYTES, &tmap_a, &full_bar[s], kb * BK, (int32_t)m0);
                probe_tma_load_2d_pair(smem_b + s * B_BYTES, &tmap_b, &full_bar[s], kb * BK, (int32_t)(n0 + rank * BNH));
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc_mxf4(2 * BM, BN);
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait_bounded(&full_bar[s], ph, 200 + s);
                tc_fence_after();
                const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + s * A_BYTES));
                const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + s * B_BYTES));
#pragma unroll
                for (int k = 0; k < BK / 32; ++k)
                    probe_mma_mxf4_pair(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, tmem_base + SF_COL,
                                  tmem_base + SF_COL, (kb | k) != 0);
                probe_tc_commit_pair(&empty_bar[s]);
            }
            probe_tc_commit_pair(accum_bar);
        }
    } else {
        const int q = warp & 3;
        const int64_t row = m0 + q * 32 + lane;
        mbar_wait_bounded(accum_bar, 0, 300);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
            if (row < M) {
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if (n0 + c0 + e < N) D[row * ldd + n0 + c0 + e] = __float2int_rn(__uint_as_float(v[e]));
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    probe_cluster_sync();                                // neither CTA leaves (or frees TMEM) while the pair is still in flight
    if (warp == 1) {
        tc_fence_after();
        probe_tmem_dealloc_pair(tmem_base);
    }
}

// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_tmap(const void *base, uint64_t row_bytes, uint64_t rows, uint32_t box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
            fprintf(stderr, "cuTensorMapEncodeTiled not available\n");
            exit(2);
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    CUtensorMap m;
    cuuint64_t dims[2] = {row_bytes, rows};
    cuuint64_t strides[1] = {row_bytes};
    cuuint32_t box[2] = {128, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        fprintf(stderr, "cuTensorMapEncodeTiled failed (%d)\n", (int)rc);
        exit(2);
    }
    return m;
}

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s failed at line %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

// nibble codes of the values 0, 1, 2 in e2m1
static const uint8_t kCode[3] = {0x0, 0x2, 0x4};

static void fill(std::vector<uint8_t> &bytes, std::vector<uint8_t> &vals, int64_t rows, int64_t kbytes, unsigned seed, int maxv) {
    bytes.resize((size_t)rows * kbytes);
    vals.resize((size_t)rows * kbytes * 2);
    uint32_t s = seed;
    for (size_t i = 0; i < vals.size(); ++i) {
        s = s * 1664525u + 1013904223u;
        vals[i] = (uint8_t)((s >> 24) % (maxv + 1));
    }
    for (size_t b = 0; b < bytes.size(); ++b) bytes[b] = (uint8_t)(kCode[vals[2 * b]] | (kCode[vals[2 * b + 1]] << 4));   // low nibble = even index
}

static double run(int64_t M, int64_t N, int64_t kbytes, bool check, int reps) {
    std::vector<uint8_t> ha, hb, va, vb;
    fill(ha, va, M, kbytes, 1u, 1);      // A entries 0/1 (the U operand)
    fill(hb, vb, N, kbytes, 2u, 2);      // B entries 0/1/2 (the Wd operand)
    uint8_t *da, *db;
    int32_t *dd;
    const int64_t ldd = (N + 127) / 128 * 128;
    CK(cudaMalloc(&da, ha.size()));
    CK(cudaMalloc(&db, hb.size()));
    CK(cudaMalloc(&dd, (size_t)M * ldd * sizeof(int32_t)));
    CK(cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dd, 0xff, (size_t)M * ldd * sizeof(int32_t)));
    const CUtensorMap ta = make_tmap(da, kbytes, M, BM), tb = make_tmap(db, kbytes, N, BNH);
    CK(cudaFuncSetAttribute(dist_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int tiles_x = (int)((N + BN - 1) / BN), tiles_y = (int)((M + 2 * BM - 1) / (2 * BM));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        dist_pair_kernel<<<2 * tiles_x * tiles_y, THREADS, SMEM_BYTES>>>(ta, tb, (int)(kbytes / BK), dd, M, N, ldd, tiles_x);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    if (check) {
        std::vector<int32_t> hd((size_t)M * ldd);
        CK(cudaMemcpy(hd.data(), dd, hd.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        long long bad = 0;
        for (int64_t i = 0; i < M; ++i)
            for (int64_t j = 0; j < N; ++j) {
                int32_t ref = 0;
                for (int64_t k = 0; k < 2 * kbytes; ++k) ref += (int32_t)va[i * 2 * kbytes + k] * (int32_t)vb[j * 2 * kbytes + k];
                if (hd[i * ldd + j] != ref && bad++ < 10)
                    printf("  mismatch D[%lld][%lld] = %d, expected %d\n", (long long)i, (long long)j, hd[i * ldd + j], ref);
            }
        printf("check %lld x %lld x %lld nibbles: %s (%lld mismatches)\n", (long long)M, (long long)N, (long long)(2 * kbytes),
               bad ? "FAILED" : "exact", bad);
        if (bad) exit(3);
    }
    cudaFree(da);
    cudaFree(db);
    cudaFree(dd);
    return best;
}

int main(int argc, char **argv) {
    int major = 0;
    CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, 0));
    if (major != 10) {
        fprintf(stderr, "needs an sm_100 GPU\n");
        return 2;
    }
    // ragged small shape first: full and partial tiles in both directions, 3 K slabs
    run(300, 700, 384, true, 1);
    run(512, 512, 1024, true, 1);
    const int64_t M = argc > 1 ? atoll(argv[1]) : 16384, N = argc > 2 ? atoll(argv[2]) : 16384;
    const int64_t kbytes = argc > 3 ? atoll(argv[3]) : 51200;     // 102 400 nibbles: 51 200 genotype columns
    const double ms = run(M, N, kbytes, false, 5);
    const double ops = 2.0 * (double)M * (double)N * (2.0 * (double)kbytes);
    printf("pair kernel %lld x %lld x %lld nibbles: %.3f ms, %.2f POP/s (single-CTA tc_dist on C5's first pass: 6.2)\n",
           (long long)M, (long long)N, (long long)(2 * kbytes), ms, ops / ms / 1e12);
    return 0;
}
