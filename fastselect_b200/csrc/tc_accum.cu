// tc_accum.cu -- per-feature hit/miss weight accumulation for discrete columns as a
// (neighbour mask) x (one-hot) int8 GEMM on the 5th-gen tensor cores, with a fused
// select / scale / reduce epilogue.
//
// Replaces the discrete branch of the reference's accumulation loops and their
// normalisation (MultiSURF.py:218-251, SURF.py:165-195).  For a discrete feature f with
// value codes c_jf:
//     sum_j c_ij [x_if != x_jf] = sum_j c_ij - sum_j c_ij [c_jf = c_if],
// and with c_ij = -aH_i * mH_ij + aM_i * mM_ij (mH, mM in {-1, 0, 1}: near/far hit and
// miss masks, aH_i = 1/|H_i|, aM_i = 1/|M_i| or 1) the inner sums are the integer GEMMs
//     GH[(f,v), i] = sum_j At[(f,v), j] * mH[i, j],   GM likewise,
// over the REDUCED one-hot rows v < last_f (onehot.cu); the implied last plane follows from
// sum_v G[(f,v), i] = rs_i (the mask's row sum), so
//     W_i[f] = sum_{v<last} [c_if = v] * k_i (rs_i - G[(f,v), i])  +  [c_if = last] * k_i * sum_{v<last} G[(f,v), i]
// for each of the two masks (k_i = -aH_i or +aM_i).  The thread that owns one-hot row (f,v)
// therefore adds k_i (rs_i - G) for targets whose code is v and k_i G for targets whose
// code is the last one: no exchange between rows.  All counts are exact integers; the
// per-target scaling and the reduction over targets are done in float64.
//
// Kernel: a CTA owns 128 one-hot rows (UMMA M = TMEM lanes) and a group of up to 8
// tiles of 256 target rows (UMMA N).  Each tile is processed as two work items, "hit"
// and "miss", one mask and one 256-column TMEM accumulator each (double-buffered, so the
// epilogue of one item overlaps the MMAs of the next).  Warp 0: TMA producer (At tile +
// mask tile per K block of 128 samples, 128B swizzle, 4-stage mbarrier ring); warp 1:
// one thread issues tcgen05.mma.cta_group::1.kind::i8 (M=128, N=256, K=32); warps 2-9
// (lane quarter x column half): epilogue -- tcgen05.ld the accumulator, test the target's
// own value code (codesT), scale, and add into float64 registers per one-hot row;
// the reduction over a tile's targets is a loop over TMEM columns inside one thread.
// Samples are class-sorted, so for a class-homogeneous target tile the hit mask is
// non-zero only in the K blocks of the tile's own class and the miss mask only outside:
// the other K blocks are skipped, which keeps the MMA work at 2 MAC per (pair, feature)
// for 3-valued genotypes.  Partials are written per (tile group, one-hot row) and reduced
// in a fixed order.
#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace fs {

namespace {
constexpr int BM = 128;   // one-hot rows per CTA
constexpr int BN = 256;   // target rows per tile
constexpr int BK = 128;   // samples (bytes) per K block
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK;                  // 16 KB
constexpr int B_BYTES = BN * BK;                  // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // 48 KB
constexpr int GROUP = 8;                          // target tiles per CTA
constexpr int CONST_BYTES = 2 * BN * 16;          // [2][BN] x {double c; int rs; int pad}
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 512 + CONST_BYTES;
constexpr int THREADS = 320;              // TMA warp, MMA warp, 8 epilogue warps
constexpr int HALF = BN / 2;              // target columns per epilogue thread
constexpr int TMEM_COLS = 512;                    // 2 accumulator buffers x 256 columns

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
// does work item (tile, phase) need K block kb?  phase 0 = hit mask, 1 = miss mask
__device__ __forceinline__ bool need_block(const TileFlags &f, int phase, int kb) {
    const int64_t k0 = (int64_t)kb * BK;
    if (phase == 0) return k0 < f.he && k0 + BK > f.hs;
    return f.mixed || !(k0 >= f.hs && k0 + BK <= f.he);
}
__device__ __forceinline__ int count_blocks(const TileFlags &f, int phase, int num_k_blocks) {
    int c = 0;
    for (int kb = 0; kb < num_k_blocks; ++kb) c += need_block(f, phase, kb) ? 1 : 0;
    return c;
}
}  // namespace

__global__ void __launch_bounds__(THREADS, 1)
tc_accum_kernel(const __grid_constant__ CUtensorMap tmap_at, const __grid_constant__ CUtensorMap tmap_mh,
                const __grid_constant__ CUtensorMap tmap_mm, int num_k_blocks, int64_t n, int64_t R,
                int num_tiles, const int64_t *__restrict__ ids, int contiguous, const int32_t *__restrict__ y,
                const int64_t *__restrict__ cls_start, const RowInfo *__restrict__ rinfo,
                const uint8_t *__restrict__ codesT, int64_t ldt, const uint32_t *__restrict__ krow, int64_t K_rows,
                double *__restrict__ tpartial) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tfull_bar = empty_bar + STAGES;     // [2] accumulator buffer ready
    uint64_t *tempty_bar = tfull_bar + 2;         // [2] accumulator buffer drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);
    // per-target constants of the work item being drained: [2][BN] x {double c; int rs}
    unsigned char *s_const = smem + STAGES * STAGE_BYTES + 512;

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
            tc::mbar_init(&tempty_bar[b], 256);           // all epilogue threads arrive
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
                for (int phase = 0; phase < 2; ++phase) {
                    const CUtensorMap *tm = phase == 0 ? &tmap_mh : &tmap_mm;
                    for (int kb = 0; kb < num_k_blocks; ++kb) {
                        if (!need_block(tf, phase, kb)) continue;
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        ++it;
                        tc::mbar_wait(&empty_bar[s], ph ^ 1);
                        unsigned char *st = smem + s * STAGE_BYTES;
                        tc::mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                        tc::tma_load_2d(st, &tmap_at, &full_bar[s], kb * BK, m0);
                        tc::tma_load_2d(st + A_BYTES, tm, &full_bar[s], kb * BK, t * BN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = tc::make_idesc_i8(BM, BN);
            int it = 0, item = 0;
            for (int t = tile_begin; t < tile_end; ++t) {
                const TileFlags tf = tile_flags(ids, R, t, contiguous != 0, y, cls_start, n);
                for (int phase = 0; phase < 2; ++phase) {
                    if (count_blocks(tf, phase, num_k_blocks) == 0) continue;   // mask is all zero: nothing to add
                    const int buf = item & 1;
                    const uint32_t tph = (item >> 1) & 1;
                    ++item;
                    tc::mbar_wait(&tempty_bar[buf], tph ^ 1);     // epilogue has drained this buffer
                    tc::tc_fence_after();
                    const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
                    uint32_t have = 0;
                    for (int kb = 0; kb < num_k_blocks; ++kb) {
                        if (!need_block(tf, phase, kb)) continue;
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        ++it;
                        tc::mbar_wait(&full_bar[s], ph);
                        tc::tc_fence_after();
                        const uint32_t sa = tc::smem_u32(smem + s * STAGE_BYTES);
                        const uint64_t da = tc::make_smem_desc_sw128(sa);
                        const uint64_t db = tc::make_smem_desc_sw128(sa + A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 32; ++k) {
                            tc::mma_i8(acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, have);
                            have = 1;
                        }
                        tc::tc_commit(&empty_bar[s]);
                    }
                    tc::tc_commit(&tfull_bar[buf]);
                }
            }
        }
    } else {
        // ===== epilogue (warps 2..9: lane quarter = warp % 4, column half = (warp - 2) / 4) =====
        // thread = one-hot row (TMEM lane) x half of the work item's 256 target columns
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int et = q * 32 + lane;                         // 0..127: TMEM lane / local one-hot row
        const int ethread = (warp - 2) * 32 + lane;           // 0..255
        const int64_t mrow = (int64_t)m0 + et;
        const bool row_live = mrow < K_rows;
        // this one-hot row's column (codesT row), value code and the column's last code
        const uint32_t meta = row_live ? krow[mrow] : 0u;
        const uint8_t *at_row = codesT + (int64_t)(meta & 0xffffffu) * ldt;
        const uint32_t own4 = ((meta >> 24) & 0xfu) * 0x01010101u;
        // dead rows (beyond K_rows) match nothing: 0xff is not a code
        const uint32_t last4 = row_live ? (meta >> 28) * 0x01010101u : 0xffffffffu;
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        int item = 0;
        for (int t = tile_begin; t < tile_end; ++t) {
            const TileFlags tf = tile_flags(ids, R, t, contiguous != 0, y, cls_start, n);
            // value codes of this thread's 128 targets in its column (issued before the
            // accumulator wait so the loads overlap the MMAs); shared by both phases
            uint32_t oh[HALF / 4];
            {
                const int64_t rbase = (int64_t)t * BN + half * HALF;
                const int64_t id0 = contiguous ? ids[0] + rbase : 0;
                if (contiguous && rbase + HALF <= R && (id0 & 15) == 0) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(at_row + id0);
#pragma unroll
                    for (int w = 0; w < HALF / 16; ++w) {
                        const uint4 a = src[w];
                        oh[4 * w] = a.x; oh[4 * w + 1] = a.y; oh[4 * w + 2] = a.z; oh[4 * w + 3] = a.w;
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < HALF / 4; ++w) {
                        uint32_t x = 0;
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int64_t r = rbase + w * 4 + b;
                            uint32_t byte = 0xffu;      // not a code: contributes nothing
                            if (r < R) byte = (uint32_t)(contiguous ? at_row[id0 + w * 4 + b] : at_row[ids[r]]);
                            x |= byte << (8 * b);
                        }
                        oh[w] = x;
                    }
                }
            }
            for (int phase = 0; phase < 2; ++phase) {
                if (count_blocks(tf, phase, num_k_blocks) == 0) continue;
                const int buf = item & 1;
                const uint32_t tph = (item >> 1) & 1;
                ++item;
                // per-target constants: c = -aH (hit phase) or +aM (miss phase); rs = mask row sum
                double *s_c = reinterpret_cast<double *>(s_const + buf * BN * 16);
                int32_t *s_rs = reinterpret_cast<int32_t *>(s_const + buf * BN * 16 + BN * 8);
                {
                    const int64_t r = (int64_t)t * BN + ethread;
                    double c = 0.0;
                    int rs = 0;
                    if (r < R) {
                        const RowInfo ri = rinfo[r];
                        c = phase == 0 ? ri.coef[FS_MASK_NEAR_HIT] : ri.coef[FS_MASK_NEAR_MISS];
                        rs = phase == 0 ? ri.n_hit - ri.n_far_hit : ri.n_miss - ri.n_far_miss;
                    }
                    s_c[ethread] = c;
                    s_rs[ethread] = rs;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                tc::mbar_wait(&tfull_bar[buf], tph);
                tc::tc_fence_after();
                const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * HALF);
#pragma unroll
                for (int c0 = 0; c0 < HALF; c0 += 32) {
                    uint32_t v[32];
                    tc::tmem_ld_32x32(tacc + c0, v);
                    tc::tmem_ld_wait();
                    const double *cc = s_c + half * HALF + c0;
                    const int32_t *rr = s_rs + half * HALF + c0;
#pragma unroll
                    for (int e = 0; e < 32; e += 4) {
                        const uint32_t w = oh[(c0 + e) >> 2];
                        const uint32_t own = __vcmpeq4(w, own4), lst = __vcmpeq4(w, last4);
                        // t = own ? rs - G : (last ? G : 0); four independent float64 chains
                        const int g0 = (int)v[e], g1 = (int)v[e + 1], g2 = (int)v[e + 2], g3 = (int)v[e + 3];
                        const int t0 = (own & 0x000000ffu) ? rr[e] - g0 : ((lst & 0x000000ffu) ? g0 : 0);
                        const int t1 = (own & 0x0000ff00u) ? rr[e + 1] - g1 : ((lst & 0x0000ff00u) ? g1 : 0);
                        const int t2 = (own & 0x00ff0000u) ? rr[e + 2] - g2 : ((lst & 0x00ff0000u) ? g2 : 0);
                        const int t3 = (own & 0xff000000u) ? rr[e + 3] - g3 : ((lst & 0xff000000u) ? g3 : 0);
                        acc0 = fma(cc[e], (double)t0, acc0);
                        acc1 = fma(cc[e + 1], (double)t1, acc1);
                        acc2 = fma(cc[e + 2], (double)t2, acc2);
                        acc3 = fma(cc[e + 3], (double)t3, acc3);
                    }
                }
                tc::tc_fence_before();
                tc::mbar_arrive(&tempty_bar[buf]);
            }
        }
        // two column halves per one-hot row: partial layout [group][half][row]
        if (row_live)
            tpartial[((int64_t)blockIdx.x * 2 + half) * K_rows + mrow] = (acc0 + acc1) + (acc2 + acc3);
    }
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

// partial vectors written per launch: (tile groups) x (2 column halves)
int tc_accum_groups(int64_t R) { return 2 * (int)ceil_div(ceil_div(R, BN), GROUP); }

void launch_tc_accum(const CUtensorMap &tmap_at, const CUtensorMap &tmap_mh, const CUtensorMap &tmap_mm, int64_t n,
                     int64_t R, const int64_t *d_ids, bool contiguous, const int32_t *d_y,
                     const int64_t *d_cls_start, const RowInfo *rinfo, const uint8_t *codesT, int64_t ldt,
                     const uint32_t *krow, int64_t K_rows, double *tpartial, cudaStream_t st, int *launches,
                     const int64_t *h_ids, const int32_t *h_y, const int64_t *h_cls_start, double *ops) {
    FS_CUDA(cudaFuncSetAttribute(tc_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int num_tiles = (int)ceil_div(R, BN);
    dim3 grid((unsigned)(tc_accum_groups(R) / 2), (unsigned)ceil_div(K_rows, BM));
    tc_accum_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(tmap_at, tmap_mh, tmap_mm, (int)ceil_div(n, BK), n, R, num_tiles,
                                                       d_ids, contiguous ? 1 : 0, d_y, d_cls_start, rinfo, codesT,
                                                       ldt, krow, K_rows, tpartial);
    FS_CUDA(cudaGetLastError());
    ++*launches;
    if (ops) {
        // K blocks actually contracted: the kernel's own tile_flags / need_block rule, on the host
        const int nkb = (int)ceil_div(n, BK);
        int64_t blocks = 0;
        for (int t = 0; t < num_tiles; ++t) {
            int64_t hs = 0, he = n;
            bool mixed = true;
            if (contiguous) {
                const int64_t ra = h_ids[0] + (int64_t)t * BN;
                const int64_t rb = std::min<int64_t>(ra + BN, h_ids[0] + R);
                const int c_lo = h_y[ra], c_hi = h_y[rb - 1];
                hs = h_cls_start[c_lo];
                he = h_cls_start[c_hi + 1];
                mixed = c_lo != c_hi;
            }
            for (int kb = 0; kb < nkb; ++kb) {
                const int64_t k0 = (int64_t)kb * BK;
                blocks += (k0 < he && k0 + BK > hs) ? 1 : 0;
                blocks += (mixed || !(k0 >= hs && k0 + BK <= he)) ? 1 : 0;
            }
        }
        *ops += 2.0 * BM * BN * BK * (double)blocks * (double)grid.y;
    }
}

}  // namespace fs
