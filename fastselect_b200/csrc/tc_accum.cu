// tc_accum.cu -- per-feature hit/miss weight accumulation for discrete columns as a
// (neighbour mask) x (one-hot) int8 GEMM on the 5th-gen tensor cores, with a fused
// select / scale / reduce epilogue.
//
// Replaces the discrete branch of the reference's accumulation loops and their
// normalisation (MultiSURF.py:218-251, SURF.py:165-195).  For a discrete feature f with
// one-hot rows (f, v):
//     sum_j c_ij [x_if != x_jf] = sum_j c_ij - sum_j c_ij A_{x_if}[j, f],
// and with c_ij = -aH_i * mH_ij + aM_i * mM_ij (mH, mM in {-1, 0, 1}: near/far hit and
// miss masks, aH_i = 1/|H_i|, aM_i = 1/|M_i| or 1) the inner sums are the integer GEMMs
//     GH[(f,v), i] = sum_j At[(f,v), j] * mH[i, j],   GM likewise,
// so  W_i[f] = sum_v At[(f,v), i] * ( -aH_i (rsH_i - GH) + aM_i (rsM_i - GM) ),
// rsH_i / rsM_i being the row sums of the masks.  All counts are exact integers; the
// per-target scaling and the reduction over targets are done in float64.
//
// Kernel: a CTA owns 128 one-hot rows (UMMA M = TMEM lanes) and a group of up to 8
// tiles of 128 target rows (UMMA N).  Warp 0: TMA producer (At tile + the mask tiles a
// K block needs); warp 1: one thread issues tcgen05.mma kind::i8 into the hit and miss
// accumulators (2 x 128 TMEM columns, double-buffered across target tiles); warps 2-5:
// epilogue -- tcgen05.ld both accumulators, pick the plane the target itself carries
// (the one-hot byte At[(f,v), i]), scale, and add into one float64 register per one-hot
// row; the reduction over a tile's targets is a loop over TMEM columns inside one
// thread.  Samples are class-sorted, so for a class-homogeneous target tile the hit
// mask is non-zero only in the K blocks of the tile's own class and the miss mask only
// outside: the other K blocks are skipped, which keeps the MMA work at 3 MAC per
// (pair, feature).  Partials are written per (tile group, one-hot row) and reduced in
// a fixed order.
#include "common.cuh"
#include "tc_common.cuh"

namespace fs {

namespace {
constexpr int BM = 128;   // one-hot rows per CTA
constexpr int BN = 128;   // target rows per tile
constexpr int BK = 128;   // samples (bytes) per K block
constexpr int STAGES = 4;
constexpr int TILE_BYTES = 128 * BK;              // 16 KB (A, mH and mM tiles alike)
constexpr int STAGE_BYTES = 3 * TILE_BYTES;       // 48 KB
constexpr int GROUP = 8;                          // target tiles per CTA
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 512 + 2 * BN * 24;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;                    // 2 buffers x (hit, miss) x 128 columns

struct TileFlags {
    int64_t hs, he;   // sample range that can hold hits of this target tile
    bool mixed;       // tile spans more than one class (or targets are not contiguous)
};

__device__ __forceinline__ TileFlags tile_flags(const int64_t *ids, int64_t R, int tile, bool contiguous,
                                                const int32_t *y, const int64_t *cls_start, int64_t n) {
    TileFlags f;
    if (!contiguous) {
        f.hs = 0; f.he = n; f.mixed = true;
        return f;
    }
    const int64_t ra = ids[0] + (int64_t)tile * BN;
    int64_t rb = ra + BN;
    const int64_t rend = ids[0] + R;
    rb = rb < rend ? rb : rend;
    const int c_lo = y[ra], c_hi = y[rb - 1];
    f.hs = cls_start[c_lo];
    f.he = cls_start[c_hi + 1];
    f.mixed = c_lo != c_hi;
    return f;
}
__device__ __forceinline__ bool need_hit(const TileFlags &f, int kb) {
    const int64_t k0 = (int64_t)kb * BK;
    return k0 < f.he && k0 + BK > f.hs;
}
__device__ __forceinline__ bool need_miss(const TileFlags &f, int kb) {
    const int64_t k0 = (int64_t)kb * BK;
    return f.mixed || !(k0 >= f.hs && k0 + BK <= f.he);
}
}  // namespace

__global__ void __launch_bounds__(THREADS, 1)
tc_accum_kernel(const __grid_constant__ CUtensorMap tmap_at, const __grid_constant__ CUtensorMap tmap_mh,
                const __grid_constant__ CUtensorMap tmap_mm, int num_k_blocks, int64_t n, int64_t R,
                int num_tiles, const int64_t *__restrict__ ids, int contiguous, const int32_t *__restrict__ y,
                const int64_t *__restrict__ cls_start, const RowInfo *__restrict__ rinfo,
                const int8_t *__restrict__ At, int64_t ldt, int64_t K_rows, double *__restrict__ tpartial) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tfull_bar = empty_bar + STAGES;     // [2] accumulator buffer ready
    uint64_t *tempty_bar = tfull_bar + 2;         // [2] accumulator buffer drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);
    // per-target constants of the tile being drained: [2][BN] x {cH, cM (double), rsH, rsM (int)}
    double *s_c = reinterpret_cast<double *>(smem + STAGES * STAGE_BYTES + 512);
    int32_t *s_rs = reinterpret_cast<int32_t *>(s_c + 2 * BN * 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM;                       // first one-hot row of this CTA
    const int tile_begin = blockIdx.x * GROUP;
    const int tile_end = tile_begin + GROUP < num_tiles ? tile_begin + GROUP : num_tiles;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_at);
        tc::prefetch_tmap(&tmap_mh);
        tc::prefetch_tmap(&tmap_mm);
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&tfull_bar[b], 1);
            tc::mbar_init(&tempty_bar[b], 128);           // all epilogue threads arrive
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;
            for (int t = tile_begin; t < tile_end; ++t) {
                const TileFlags tf = tile_flags(ids, R, t, contiguous != 0, y, cls_start, n);
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    const bool nh = need_hit(tf, kb), nm = need_miss(tf, kb);
                    if (!nh && !nm) continue;
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    ++it;
                    tc::mbar_wait(&empty_bar[s], ph ^ 1);
                    unsigned char *st = smem + s * STAGE_BYTES;
                    tc::mbar_arrive_expect_tx(&full_bar[s], TILE_BYTES * (1 + (nh ? 1 : 0) + (nm ? 1 : 0)));
                    tc::tma_load_2d(st, &tmap_at, &full_bar[s], kb * BK, m0);
                    if (nh) tc::tma_load_2d(st + TILE_BYTES, &tmap_mh, &full_bar[s], kb * BK, t * BN);
                    if (nm) tc::tma_load_2d(st + 2 * TILE_BYTES, &tmap_mm, &full_bar[s], kb * BK, t * BN);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = tc::make_idesc_i8(BM, BN);
            int it = 0;
            for (int t = tile_begin; t < tile_end; ++t) {
                const int buf = (t - tile_begin) & 1;
                const uint32_t tph = ((t - tile_begin) >> 1) & 1;
                const TileFlags tf = tile_flags(ids, R, t, contiguous != 0, y, cls_start, n);
                tc::mbar_wait(&tempty_bar[buf], tph ^ 1);     // epilogue has drained this buffer
                tc::tc_fence_after();
                const uint32_t acc_h = tmem_base + (uint32_t)(buf * 2 * BN);
                const uint32_t acc_m = acc_h + BN;
                uint32_t have_h = 0, have_m = 0;
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    const bool nh = need_hit(tf, kb), nm = need_miss(tf, kb);
                    if (!nh && !nm) continue;
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    ++it;
                    tc::mbar_wait(&full_bar[s], ph);
                    tc::tc_fence_after();
                    const uint32_t sa = tc::smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t da = tc::make_smem_desc_sw128(sa);
                    const uint64_t dh = tc::make_smem_desc_sw128(sa + TILE_BYTES);
                    const uint64_t dm = tc::make_smem_desc_sw128(sa + 2 * TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 32; ++k) {
                        if (nh) { tc::mma_i8(acc_h, da + (uint64_t)(2 * k), dh + (uint64_t)(2 * k), idesc, have_h); have_h = 1; }
                        if (nm) { tc::mma_i8(acc_m, da + (uint64_t)(2 * k), dm + (uint64_t)(2 * k), idesc, have_m); have_m = 1; }
                    }
                    tc::tc_commit(&empty_bar[s]);
                }
                tc::tc_commit(&tfull_bar[buf]);
            }
        }
    } else {
        // ===== epilogue (warps 2..5, 128 threads, thread = one-hot row) =====
        const int q = warp & 3;
        const int et = q * 32 + lane;                         // 0..127: TMEM lane / local one-hot row
        const int64_t mrow = (int64_t)m0 + et;
        const bool row_live = mrow < K_rows;
        const int8_t *at_row = At + (row_live ? mrow : 0) * ldt;
        double acc = 0.0;
        for (int t = tile_begin; t < tile_end; ++t) {
            const int buf = (t - tile_begin) & 1;
            const uint32_t tph = ((t - tile_begin) >> 1) & 1;
            const TileFlags tf = tile_flags(ids, R, t, contiguous != 0, y, cls_start, n);
            // does either accumulator receive no MMA at all for this tile?
            bool any_h = false, any_m = false;
            for (int kb = 0; kb < num_k_blocks; ++kb) { any_h |= need_hit(tf, kb); any_m |= need_miss(tf, kb); }
            // per-target constants (one target per epilogue thread)
            {
                const int64_t r = (int64_t)t * BN + et;
                double ch = 0.0, cm = 0.0;
                int rh = 0, rm = 0;
                if (r < R) {
                    const RowInfo ri = rinfo[r];
                    ch = ri.coef[FS_MASK_NEAR_HIT];           // -aH
                    cm = ri.coef[FS_MASK_NEAR_MISS];          // +aM
                    rh = ri.n_hit - ri.n_far_hit;
                    rm = ri.n_miss - ri.n_far_miss;
                }
                s_c[(buf * BN + et) * 2] = ch;
                s_c[(buf * BN + et) * 2 + 1] = cm;
                s_rs[(buf * BN + et) * 2] = rh;
                s_rs[(buf * BN + et) * 2 + 1] = rm;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            tc::mbar_wait(&tfull_bar[buf], tph);
            tc::tc_fence_after();
            const uint32_t acc_h = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * BN);
            const uint32_t acc_m = acc_h + BN;
            const int64_t id0 = contiguous ? ids[0] + (int64_t)t * BN : 0;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t vh[32], vm[32];
                if (any_h) tc::tmem_ld_32x32(acc_h + c0, vh);
                if (any_m) tc::tmem_ld_32x32(acc_m + c0, vm);
                // one-hot bytes of the 32 targets at this one-hot row
                uint32_t oh[8];
                if (contiguous) {
                    // ids are contiguous and tile-aligned to 128 relative to ids[0]; ids[0] may be unaligned
                    const int8_t *src = at_row + id0 + c0;
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        uint32_t x = 0;
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int64_t r = (int64_t)t * BN + c0 + w * 4 + b;
                            const uint32_t byte = (r < R) ? (uint32_t)(uint8_t)src[w * 4 + b] : 0u;
                            x |= byte << (8 * b);
                        }
                        oh[w] = x;
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        uint32_t x = 0;
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int64_t r = (int64_t)t * BN + c0 + w * 4 + b;
                            const uint32_t byte = (r < R) ? (uint32_t)(uint8_t)at_row[ids[r]] : 0u;
                            x |= byte << (8 * b);
                        }
                        oh[w] = x;
                    }
                }
                tc::tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const bool on = ((oh[e >> 2] >> (8 * (e & 3))) & 0xffu) != 0u;
                    const int c = c0 + e;
                    const double ch = s_c[(buf * BN + c) * 2], cm = s_c[(buf * BN + c) * 2 + 1];
                    const int gh = any_h ? (int)vh[e] : 0, gm = any_m ? (int)vm[e] : 0;
                    const double term = ch * (double)(s_rs[(buf * BN + c) * 2] - gh) +
                                        cm * (double)(s_rs[(buf * BN + c) * 2 + 1] - gm);
                    acc += on ? term : 0.0;
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(&tempty_bar[buf]);
        }
        if (row_live) tpartial[(int64_t)blockIdx.x * K_rows + mrow] = acc;
    }
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

int tc_accum_groups(int64_t R) { return (int)ceil_div(ceil_div(R, BN), GROUP); }

void launch_tc_accum(const CUtensorMap &tmap_at, const CUtensorMap &tmap_mh, const CUtensorMap &tmap_mm, int64_t n,
                     int64_t R, const int64_t *d_ids, bool contiguous, const int32_t *d_y,
                     const int64_t *d_cls_start, const RowInfo *rinfo, const int8_t *At, int64_t ldt, int64_t K_rows,
                     double *tpartial, cudaStream_t st, int *launches) {
    FS_CUDA(cudaFuncSetAttribute(tc_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int num_tiles = (int)ceil_div(R, BN);
    dim3 grid((unsigned)tc_accum_groups(R), (unsigned)ceil_div(K_rows, BM));
    tc_accum_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(tmap_at, tmap_mh, tmap_mm, (int)ceil_div(n, BK), n, R, num_tiles,
                                                       d_ids, contiguous ? 1 : 0, d_y, d_cls_start, rinfo, At, ldt,
                                                       K_rows, tpartial);
    FS_CUDA(cudaGetLastError());
    ++*launches;
}

}  // namespace fs
