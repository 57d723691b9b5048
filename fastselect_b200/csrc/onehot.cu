// onehot.cu -- one-hot tensor-core path: encoding of the active discrete columns, TMA
// tensor maps, and the host side of the tcgen05 distance / accumulation kernels.
//
// K0 "encode" of the design: every active discrete column f with V_f <= FS_DISTINCT_CAP
// distinct values becomes V_f one-hot int8 rows/columns (one per value; only equality
// matters, MultiSURF.py:184-185, so any injective value coding is exact):
//   A  [n, K]   sample-major (one-hot index contiguous)  -> operands of the distance GEMM
//   At [K, ldt] feature-major (sample index contiguous)   -> A operand of the accumulation GEMM
//   codes [n, ldc] the value index itself (ReliefF's sparse neighbour gather)
// Samples are in the data set's class-sorted internal order.  HBM-bound streaming kernel:
// reads n*pt elements once, writes n*K*2 + n*pt bytes.
#include <algorithm>

#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace fs {

void launch_tc_dist(const CUtensorMap &tmap_a, const CUtensorMap &tmap_b, int64_t K, int32_t p_disc, int64_t R,
                    int64_t n, int32_t *Dd, int64_t ldd, cudaStream_t st, int *launches);
int tc_accum_groups(int64_t R);
void launch_tc_accum(const CUtensorMap &tmap_at, const CUtensorMap &tmap_mh, const CUtensorMap &tmap_mm, int64_t n,
                     int64_t R, const int64_t *d_ids, bool contiguous, const int32_t *d_y,
                     const int64_t *d_cls_start, const RowInfo *rinfo, const int8_t *At, int64_t ldt, int64_t K_rows,
                     double *tpartial, cudaStream_t st, int *launches);

bool tensor_path_available() { return true; }

// ---------------------------------------------------------------------------
// TMA descriptor (driver entry point fetched through the runtime: no -lcuda)
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        FS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        FS_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, FS_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

CUtensorMap make_tmap_u8_sw128(const void *base, uint64_t row_bytes, uint64_t rows, uint64_t pitch_bytes,
                               uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {row_bytes, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {128, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FS_REQUIRE(rc == CUDA_SUCCESS, FS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): base=%p row_bytes=%llu rows=%llu pitch=%llu",
               (int)rc, base, (unsigned long long)row_bytes, (unsigned long long)rows, (unsigned long long)pitch_bytes);
    return m;
}

// ---------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------
// Tile: 128 samples x 64 columns per CTA.  Step 1 reads x coalesced along the column
// index and keeps the value codes in shared memory (both orientations); step 2 emits the
// A rows (one-hot index contiguous) and step 3 the At rows (sample index contiguous) as
// packed 32-bit stores -- every byte of the tile's A / At footprint is written exactly
// once, so no memset of the operands is needed.
constexpr int ENC_ROWS = 128;
constexpr int ENC_COLS = 64;

template <typename Tin>
__global__ void __launch_bounds__(256) onehot_encode_kernel(const Tin *__restrict__ x, int64_t ldx,
                                                            const int64_t *__restrict__ perm,
                                                            const int64_t *__restrict__ tcol,
                                                            const int32_t *__restrict__ toff,
                                                            const double *__restrict__ vals, int as_f32, int64_t n,
                                                            int64_t pt, int64_t K, int64_t ldt, int64_t ldc,
                                                            int8_t *__restrict__ A, int8_t *__restrict__ At,
                                                            uint8_t *__restrict__ codes) {
    __shared__ __align__(16) uint8_t code_rc[ENC_ROWS][ENC_COLS];       // [sample][column]
    __shared__ __align__(16) uint8_t code_cr[ENC_COLS][ENC_ROWS + 4];   // [column][sample], padded: conflict-free transposed stores
    __shared__ uint8_t kcol[ENC_COLS * FS_DISTINCT_CAP];                // one-hot index -> column in tile
    __shared__ uint8_t kval[ENC_COLS * FS_DISTINCT_CAP];                // one-hot index -> value code
    const int tid = threadIdx.x;
    const int64_t c0 = (int64_t)blockIdx.x * ENC_COLS, r0 = (int64_t)blockIdx.y * ENC_ROWS;
    const int ncols = (int)(pt - c0 < ENC_COLS ? pt - c0 : ENC_COLS);
    const int nrows = (int)(n - r0 < ENC_ROWS ? n - r0 : ENC_ROWS);
    const int k0 = toff[c0], k1 = toff[c0 + ncols];

    // ---- step 1: value codes
    {
        const int c = tid & (ENC_COLS - 1);
        int64_t f = 0;
        int off = 0, V = 0;
        // code table of this thread's column in the input's own type (values came from the
        // data, so the conversion back is exact); float64 input scored in float32 arithmetic
        // is compared after narrowing, as the reference compares x.astype(float32)
        Tin v[FS_DISTINCT_CAP];
        float vf[FS_DISTINCT_CAP];
        if (c < ncols) {
            f = tcol[c0 + c];
            off = toff[c0 + c];
            V = toff[c0 + c + 1] - off;
#pragma unroll
            for (int q = 0; q < FS_DISTINCT_CAP; ++q) {
                const double d = vals[f * FS_DISTINCT_CAP + q];
                v[q] = (Tin)d;
                vf[q] = (float)d;
            }
            if (tid < ENC_COLS)
                for (int q = 0; q < V; ++q) {
                    kcol[off - k0 + q] = (uint8_t)c;
                    kval[off - k0 + q] = (uint8_t)q;
                }
        }
        for (int rr = tid >> 6; rr < ENC_ROWS; rr += 4) {
            uint8_t code = 0;
            if (c < ncols && rr < nrows) {
                const Tin xv = x[perm[r0 + rr] * ldx + f];
                int found = 0;
                if (V <= 4) {                        // genotype-like columns: 4 compares, no loop
#pragma unroll
                    for (int q = 3; q >= 0; --q) {
                        const bool eq = as_f32 ? ((float)xv == vf[q]) : (xv == v[q]);
                        if (q < V && eq) found = q;  // lowest matching index wins
                    }
                } else {
#pragma unroll
                    for (int q = FS_DISTINCT_CAP - 1; q >= 0; --q) {
                        const bool eq = as_f32 ? ((float)xv == vf[q]) : (xv == v[q]);
                        if (q < V && eq) found = q;
                    }
                }
                code = (uint8_t)found;
                codes[(r0 + rr) * ldc + c0 + c] = code;
            }
            code_rc[rr][c] = code;
            code_cr[c][rr] = code;
        }
    }
    __syncthreads();

    // ---- step 2: A[r, k0..k1)
    // fast path (every column of the tile has 3 values and k0 is word-aligned, i.e. 0/1/2
    // genotypes): 4 codes -> 12 one-hot bytes = three 32-bit words built with shifts
    const bool v3 = __syncthreads_and((tid >= ncols) || (toff[c0 + (tid < ncols ? tid : 0) + 1] - toff[c0 + (tid < ncols ? tid : 0)] == 3)) &&
                    (k0 & 3) == 0 && (ncols & 3) == 0;
    if (v3) {
        const int ngroups = ncols >> 2;                           // 4 columns per thread-iteration
        for (int rr = tid >> 4; rr < nrows; rr += 16)
            for (int g = tid & 15; g < ngroups; g += 16) {
                const uint32_t cw = *reinterpret_cast<const uint32_t *>(&code_rc[rr][4 * g]);
                // p_i = one-hot triple of code i as a 24-bit little-endian pattern
                const uint32_t p0 = 1u << (8 * (cw & 0xffu)), p1 = 1u << (8 * ((cw >> 8) & 0xffu));
                const uint32_t p2 = 1u << (8 * ((cw >> 16) & 0xffu)), p3 = 1u << (8 * (cw >> 24));
                uint32_t *dst = reinterpret_cast<uint32_t *>(A + (r0 + rr) * K + k0 + 12 * g);
                dst[0] = p0 | (p1 << 24);
                dst[1] = (p1 >> 8) | (p2 << 16);
                dst[2] = (p2 >> 16) | (p3 << 8);
            }
    } else {
        // general path: 32-bit words where the whole word belongs to this tile, single bytes at
        // the unaligned edges (a neighbouring tile owns the rest of that word)
        const int w0 = k0 >> 2, w1 = (k1 + 3) >> 2;               // word range covering [k0, k1)
        for (int rr = tid >> 5; rr < nrows; rr += 8)
        for (int w = w0 + (tid & 31); w < w1; w += 32) {
            uint32_t word = 0;
            bool full = true;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int k = 4 * w + b;
                if (k >= k0 && k < k1) {
                    const uint32_t on = code_rc[rr][kcol[k - k0]] == kval[k - k0] ? 1u : 0u;
                    word |= on << (8 * b);
                } else {
                    full = false;
                }
            }
            int8_t *dst = A + (r0 + rr) * K + 4 * (int64_t)w;
            if (full) {
                *reinterpret_cast<uint32_t *>(dst) = word;
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int k = 4 * w + b;
                    if (k >= k0 && k < k1) dst[b] = (int8_t)((word >> (8 * b)) & 0xffu);
                }
            }
        }
    }
    // ---- step 3: At[k, r0..r0+128): one warp instruction writes one 128-byte row segment
    {
        const int lane4 = (tid & 31) * 4;
        for (int k = k0 + (tid >> 5); k < k1; k += 8) {
            const uint32_t cw = *reinterpret_cast<const uint32_t *>(&code_cr[kcol[k - k0]][lane4]);
            const uint32_t val = kval[k - k0];
            const uint32_t word = __vcmpeq4(cw, val * 0x01010101u) & 0x01010101u;
            int8_t *dst = At + (int64_t)k * ldt + r0 + lane4;       // r0, ldt multiples of 128: aligned
            if (lane4 + 4 <= nrows) {
                *reinterpret_cast<uint32_t *>(dst) = word;
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (lane4 + b < nrows) dst[b] = (int8_t)((word >> (8 * b)) & 0xffu);
            }
        }
    }
}

void build_onehot(fs_dataset *ds, WorkSet &ws, int *launches) {
    const int64_t n = ds->n, pt = ws.pt;
    ws.h_toff.assign(pt + 1, 0);
    for (int64_t c = 0; c < pt; ++c) ws.h_toff[c + 1] = ws.h_toff[c] + ds->cnt[ws.h_tcol[c]];
    ws.K_used = ws.h_toff[pt];
    ws.K = round_up(ws.K_used, 128);
    ws.ldt = round_up(n, 128);
    ws.ldc = round_up(pt, 16);
    FS_REQUIRE(ws.K < (1LL << 31), FS_ERR_INVALID, "one-hot contraction length too large");
    std::vector<int32_t> toff32(ws.h_toff.begin(), ws.h_toff.end());
    ws.tcol.reserve(pt);
    ws.tout.reserve(pt);
    ws.toff.reserve(pt + 1);
    ws.A.reserve((size_t)n * ws.K);
    ws.At.reserve((size_t)ws.K * ws.ldt);
    ws.codes.reserve((size_t)n * ws.ldc);
    cudaStream_t st = ds->stream;
    FS_CUDA(cudaMemcpyAsync(ws.tcol.ptr, ws.h_tcol.data(), pt * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    FS_CUDA(cudaMemcpyAsync(ws.tout.ptr, ws.h_tout.data(), pt * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    FS_CUDA(cudaMemcpyAsync(ws.toff.ptr, toff32.data(), (pt + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    // the encode kernel writes every used byte of A, At and codes exactly once; only the K
    // padding (one-hot indices K_used..K) has to be cleared.  Sample padding of At rows
    // (columns n..ldt) and code padding are never read (the TMA maps are n bytes wide).
    if (ws.K > ws.K_used) {
        FS_CUDA(cudaMemset2DAsync(ws.A.ptr + ws.K_used, (size_t)ws.K, 0, (size_t)(ws.K - ws.K_used), (size_t)n, st));
        FS_CUDA(cudaMemsetAsync(ws.At.ptr + (size_t)ws.K_used * ws.ldt, 0, (size_t)(ws.K - ws.K_used) * ws.ldt, st));
    }
    dim3 grid((unsigned)ceil_div(pt, ENC_COLS), (unsigned)ceil_div(n, ENC_ROWS));
    const int as_f32 = (ds->arith == FS_ARITH_F32 && ds->dtype == FS_F64) ? 1 : 0;
#define FS_ENCODE(T)                                                                                              \
    onehot_encode_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T *>(ds->x), ds->ldx, ds->d_perm.ptr,        \
                                                  ws.tcol.ptr, ws.toff.ptr, ds->d_vals.ptr, as_f32, n, pt, ws.K, \
                                                  ws.ldt, ws.ldc, ws.A.ptr, ws.At.ptr, ws.codes.ptr)
    switch (ds->dtype) {
        case FS_U8: FS_ENCODE(uint8_t); break;
        case FS_I8: FS_ENCODE(int8_t); break;
        case FS_F32: FS_ENCODE(float); break;
        case FS_F64: FS_ENCODE(double); break;
    }
#undef FS_ENCODE
    FS_CUDA(cudaGetLastError());
    ++*launches;
    FS_CUDA(cudaStreamSynchronize(st));   // staging vectors above are stack/temporary
}

// ---------------------------------------------------------------------------
// distance
// ---------------------------------------------------------------------------
void launch_dist_tensor(fs_dataset *ds, const WorkSet &ws, int64_t r0_internal, const int64_t *h_row_ids,
                        bool contiguous, int64_t R, int32_t *Dd, int64_t ldn, cudaStream_t st, int *launches) {
    const int8_t *a_rows;
    if (contiguous) {
        a_rows = ws.A.ptr + (size_t)r0_internal * ws.K;
    } else {
        DevBuf<int8_t> &g = ds->a_gather;
        g.reserve((size_t)R * ws.K);
        for (int64_t r = 0; r < R; ++r)
            FS_CUDA(cudaMemcpyAsync(g.ptr + (size_t)r * ws.K, ws.A.ptr + (size_t)h_row_ids[r] * ws.K, ws.K,
                                    cudaMemcpyDeviceToDevice, st));
        a_rows = g.ptr;
    }
    const CUtensorMap ta = make_tmap_u8_sw128(a_rows, ws.K, R, ws.K, 128);
    const CUtensorMap tb = make_tmap_u8_sw128(ws.A.ptr, ws.K, ds->n, ws.K, 256);
    launch_tc_dist(ta, tb, ws.K, (int32_t)ws.pt, R, ds->n, Dd, ldn, st, launches);
}

// ---------------------------------------------------------------------------
// accumulation
// ---------------------------------------------------------------------------
// wsum[tout[c]] += sum over groups g and one-hot rows of column c of tpartial[g, row]
__global__ void __launch_bounds__(256) reduce_tensor_partials_kernel(const double *__restrict__ tpartial, int groups,
                                                                     int64_t K_rows, const int32_t *__restrict__ toff,
                                                                     const int64_t *__restrict__ tout, int64_t pt,
                                                                     double *__restrict__ wsum) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= pt) return;
    double s = 0.0;
    for (int g = 0; g < groups; ++g)
        for (int r = toff[c]; r < toff[c + 1]; ++r) s += tpartial[(int64_t)g * K_rows + r];
    wsum[tout[c]] += s;
}

// ReliefF on the one-hot columns: sparse neighbour lists, value codes compared directly
// (ReliefF.py:181-216; at most C*k neighbours per target) -- HBM/L2-bound gather.
__global__ void __launch_bounds__(128) relieff_codes_gather_kernel(
    const uint8_t *__restrict__ codes, int64_t ldc, int64_t pt, const int64_t *__restrict__ ids,
    const int32_t *__restrict__ nbr_idx, const double *__restrict__ nbr_w, const int32_t *__restrict__ nbr_cnt,
    int nbr_cap, int64_t R, int64_t rows_per_cta, double *__restrict__ partial) {
    __shared__ int32_t sidx[256];
    __shared__ double sw[256];
    const int tid = threadIdx.x;
    const int64_t c = (int64_t)blockIdx.y * 128 + tid;
    const bool live = c < pt;
    const int64_t cc = live ? c : pt - 1;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r_end = r_begin + rows_per_cta < R ? r_begin + rows_per_cta : R;
    double total = 0.0;
    for (int64_t row = r_begin; row < r_end; ++row) {
        const uint8_t ci = codes[ids[row] * ldc + cc];
        const int cnt = nbr_cnt[row];
        for (int s0 = 0; s0 < cnt; s0 += 256) {
            const int m = cnt - s0 < 256 ? cnt - s0 : 256;
            __syncthreads();
            for (int e = tid; e < m; e += 128) {
                sidx[e] = nbr_idx[row * nbr_cap + s0 + e];
                sw[e] = nbr_w[row * nbr_cap + s0 + e];
            }
            __syncthreads();
#pragma unroll 4
            for (int e = 0; e < m; ++e) {
                const uint8_t cj = codes[(int64_t)sidx[e] * ldc + cc];
                total += (ci != cj) ? sw[e] : 0.0;
            }
        }
    }
    if (live) partial[(int64_t)blockIdx.x * pt + c] = total;
}

__global__ void __launch_bounds__(256) reduce_code_partials_kernel(const double *__restrict__ partial, int64_t n_part,
                                                                   int64_t pt, const int64_t *__restrict__ tout,
                                                                   double *__restrict__ wsum) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= pt) return;
    double s = 0.0;
    for (int64_t q = 0; q < n_part; ++q) s += partial[q * pt + c];
    wsum[tout[c]] += s;
}

void launch_accum_tensor(fs_dataset *ds, const WorkSet &ws, int algo, const int64_t *d_row_ids,
                         const int64_t *h_row_ids, bool contiguous, int64_t R, const int8_t *sel, int64_t ldn,
                         const RowInfo *rinfo, const int32_t *nbr_idx, const double *nbr_w, const int32_t *nbr_cnt,
                         int32_t nbr_cap, double *wsum, cudaStream_t st, int *launches) {
    (void)sel;
    (void)h_row_ids;
    const int64_t n = ds->n;
    if (algo == FS_RELIEFF) {
        const int64_t ctiles = ceil_div(ws.pt, 128);
        const int64_t want = std::max<int64_t>(1, ceil_div(148 * 8, ctiles));
        const int64_t rows = std::max<int64_t>(1, ceil_div(R, want));
        const int64_t n_part = ceil_div(R, rows);
        ds->tpartial.reserve((size_t)n_part * ws.pt);
        dim3 grid((unsigned)n_part, (unsigned)ctiles);
        relieff_codes_gather_kernel<<<grid, 128, 0, st>>>(ws.codes.ptr, ws.ldc, ws.pt, d_row_ids, nbr_idx, nbr_w,
                                                          nbr_cnt, nbr_cap, R, rows, ds->tpartial.ptr);
        FS_CUDA(cudaGetLastError());
        reduce_code_partials_kernel<<<(unsigned)ceil_div(ws.pt, 256), 256, 0, st>>>(ds->tpartial.ptr, n_part, ws.pt,
                                                                                   ws.tout.ptr, wsum);
        FS_CUDA(cudaGetLastError());
        *launches += 2;
        return;
    }
    const int groups = tc_accum_groups(R);
    ds->tpartial.reserve((size_t)groups * ws.K_used);
    // K of this GEMM is the sample index: rows of At and of the masks are n bytes long
    const CUtensorMap tat = make_tmap_u8_sw128(ws.At.ptr, (uint64_t)n, (uint64_t)ws.K, (uint64_t)ws.ldt, 128);
    const CUtensorMap tmh = make_tmap_u8_sw128(ds->maskH.ptr, (uint64_t)n, (uint64_t)R, (uint64_t)ldn, 256);
    const CUtensorMap tmm = make_tmap_u8_sw128(ds->maskM.ptr, (uint64_t)n, (uint64_t)R, (uint64_t)ldn, 256);
    launch_tc_accum(tat, tmh, tmm, n, R, d_row_ids, contiguous, ds->d_y.ptr, ds->d_cls_start.ptr, rinfo, ws.At.ptr,
                    ws.ldt, ws.K_used, ds->tpartial.ptr, st, launches);
    reduce_tensor_partials_kernel<<<(unsigned)ceil_div(ws.pt, 256), 256, 0, st>>>(ds->tpartial.ptr, groups, ws.K_used,
                                                                                 ws.toff.ptr, ws.tout.ptr, ws.pt, wsum);
    FS_CUDA(cudaGetLastError());
    ++*launches;
}

}  // namespace fs
