// dist_general.cu -- CUDA-core sample-pair distance for continuous (and wide
// discrete) columns.
//
// Replaces the distance loops of the reference kernels (MultiSURF.py:181-191,
// SURF.py:151-160, ReliefF.py:149-155): d_ij = sum_f term_f(i, j) accumulated in
// float64, with term_f = fl32(fl32(|a-b|) * r_f) (FS_ARITH_F32) or |a-b| * (double)r_f
// (FS_ARITH_F64), or [a != b] for compare chunks.  Every term is computed with
// explicit round-to-nearest intrinsics (no FMA contraction) so it is bit-identical
// to the CPU oracle's; only the float64 summation order differs (O(1e-16)).
//
// Tiling: a CTA of 256 threads computes a 64 x 64 tile of D; features stream
// through shared memory in 128-byte chunks per row (32 float32 / 16 float64),
// double-buffered with cp.async.  Each thread owns a 4 x 4 register block of
// float64 accumulators (rows ty + 16a, columns tx + 16b), reads its operands as
// 128-bit shared loads from rows padded to 144 bytes (conflict-free), and the
// kernel is bound by FP32/FP64 issue, not by HBM (every loaded element is reused
// 64 times).
#include <cuda_pipeline.h>

#include "common.cuh"

namespace fs {

constexpr int kTile = 64;
constexpr int kRowBytes = kChunkBytes + 16;  // padded shared-memory row

template <typename T>
struct Vec;
template <>
struct Vec<float> {
    using type = float4;
    static constexpr int N = 4;
};
template <>
struct Vec<double> {
    using type = double2;
    static constexpr int N = 2;
};

__device__ __forceinline__ void unpack(const float4 &v, float (&o)[4]) {
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void unpack(const double2 &v, double (&o)[2]) {
    o[0] = v.x; o[1] = v.y;
}

// float32 term widened to float64 without the conversion unit: F2F.F64.F32 runs on the
// XU pipe at 16 lanes/clk/SM and bounded this kernel at 92 % XU utilisation (ncu,
// profiles/r01_ncu_general_summary.txt).  For a normal, non-negative float with bits x,
// the float64 with bits x * 2^29 + (896 << 52) has the same value (exponent re-biased from
// 127 to 1023, mantissa left-aligned): two integer instructions on the ALU/FMA pipes.  Zero (and denormal
// products, < 1.2e-38) map to ~2^-127 instead; such an addend is absorbed without trace by
// any partial sum above 2^-74, and the kernel flushes totals below 2^-100 to exactly 0, so
// only rows whose continuous distance is itself below 1e-30 could see a difference.
__device__ __forceinline__ double widen_f32(float t) {
    // written as two 32-bit halves on purpose: a single 64-bit multiply-add compiles to
    // IMAD.WIDE.U32, which stalls the FMA pipe (ncu: stall_math/dispatch on that line)
    const uint32_t x = __float_as_uint(t);
    const uint32_t hi = (x >> 3) + 0x38000000u;
    const uint32_t lo = x << 29;
    return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ double term_cont(float a, float b, float r) {
    return widen_f32(__fmul_rn(fabsf(__fsub_rn(a, b)), r));
}
__device__ __forceinline__ double term_cont(double a, double b, float r) {
    return __dmul_rn(fabs(__dsub_rn(a, b)), (double)r);
}

template <typename T>
__global__ void __launch_bounds__(256, 2)
dist_general_kernel(const T *__restrict__ xa, int64_t na, const T *__restrict__ xb, int64_t nb, int64_t ld,
                    const float *__restrict__ recip, const uint8_t *__restrict__ ctype, int nchunks,
                    double *__restrict__ D, int64_t ldd, int symmetric) {
    // symmetric: xa == xb (all samples are targets): d_ij = d_ji, so only tiles on or above
    // the diagonal are computed and each is also stored transposed (half the arithmetic)
    if (symmetric && blockIdx.x < blockIdx.y) return;
    using V = typename Vec<T>::type;
    constexpr int VN = Vec<T>::N;                 // features per 128-bit load
    constexpr int FT = kChunkBytes / sizeof(T);   // features per chunk
    constexpr int NV = FT / VN;                   // 8 vector groups per chunk

    __shared__ __align__(16) unsigned char smem[2][2][kTile * kRowBytes];  // [stage][A|B]
    __shared__ __align__(16) float srecip[2][32];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t i0 = (int64_t)blockIdx.y * kTile, j0 = (int64_t)blockIdx.x * kTile;

    // cp.async mapping: 512 x 16 B per tile, 2 per thread per tile
    auto issue = [&](int stage, int chunk) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            int idx = tid + q * 256;
            int row = idx >> 3, seg = idx & 7;
            int64_t ra = i0 + row; ra = ra < na ? ra : na - 1;
            int64_t rb = j0 + row; rb = rb < nb ? rb : nb - 1;
            const unsigned char *ga = reinterpret_cast<const unsigned char *>(xa + ra * ld + (int64_t)chunk * FT) + seg * 16;
            const unsigned char *gb = reinterpret_cast<const unsigned char *>(xb + rb * ld + (int64_t)chunk * FT) + seg * 16;
            __pipeline_memcpy_async(&smem[stage][0][row * kRowBytes + seg * 16], ga, 16);
            __pipeline_memcpy_async(&smem[stage][1][row * kRowBytes + seg * 16], gb, 16);
        }
        if (tid < 8) __pipeline_memcpy_async(&srecip[stage][tid * (FT / 8)], recip + (int64_t)chunk * FT + tid * (FT / 8), FT / 8 * 4);
        __pipeline_commit();
    };

    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

    issue(0, 0);
    for (int c = 0; c < nchunks; ++c) {
        const int st = c & 1;
        if (c + 1 < nchunks) {
            issue(st ^ 1, c + 1);
            __pipeline_wait_prior(1);
        } else {
            __pipeline_wait_prior(0);
        }
        __syncthreads();
        const unsigned char *sa = smem[st][0], *sb = smem[st][1];
        if (ctype[c] == kChunkContinuous) {
#pragma unroll 2
            for (int v = 0; v < NV; ++v) {
                T xi[4][VN], xj[4][VN];
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    unpack(*reinterpret_cast<const V *>(sa + (ty + 16 * a) * kRowBytes + v * 16), xi[a]);
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    unpack(*reinterpret_cast<const V *>(sb + (tx + 16 * b) * kRowBytes + v * 16), xj[b]);
                float r[VN];
#pragma unroll
                for (int e = 0; e < VN; ++e) r[e] = srecip[st][v * VN + e];
#pragma unroll
                for (int e = 0; e < VN; ++e)
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            acc[a][b] = __dadd_rn(acc[a][b], term_cont(xi[a][e], xj[b][e], r[e]));
            }
        } else {
            // compare chunk: mismatch counts are exact in float32 (<= 32 per chunk)
            float cnt[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) cnt[a][b] = 0.0f;
#pragma unroll 2
            for (int v = 0; v < NV; ++v) {
                T xi[4][VN], xj[4][VN];
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    unpack(*reinterpret_cast<const V *>(sa + (ty + 16 * a) * kRowBytes + v * 16), xi[a]);
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    unpack(*reinterpret_cast<const V *>(sb + (tx + 16 * b) * kRowBytes + v * 16), xj[b]);
#pragma unroll
                for (int e = 0; e < VN; ++e)
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) cnt[a][b] += (xi[a][e] != xj[b][e]) ? 1.0f : 0.0f;
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] += (double)cnt[a][b];
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = acc[a][b] < 0x1p-100 ? 0.0 : acc[a][b];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        int64_t i = i0 + ty + 16 * a;
        if (i >= na) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int64_t j = j0 + tx + 16 * b;
            if (j < nb) D[i * ldd + j] = acc[a][b];
        }
    }
    if (symmetric && blockIdx.x != blockIdx.y) {
        // transposed copy through shared memory (the operand stages are free now): 64 x 64 doubles
        double *tile = reinterpret_cast<double *>(&smem[0][0][0]);
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) tile[(tx + 16 * b) * kTile + (ty + 16 * a)] = acc[a][b];
        __syncthreads();
        for (int e = tid; e < kTile * kTile; e += 256) {
            const int jj = e >> 6, ii = e & 63;            // row of the transposed tile, column
            const int64_t j = j0 + jj, i = i0 + ii;
            if (j < nb && i < na) D[j * ldd + i] = tile[jj * kTile + ii];
        }
    }
}

void launch_dist_general(const WorkSet &ws, const void *xa, int64_t na, const void *xb, int64_t nb, double *D,
                         int64_t ldd, cudaStream_t st, int *launches) {
    if (ws.pg == 0 || na == 0) return;
    const int symmetric = (xa == xb && na == nb) ? 1 : 0;
    dim3 grid((unsigned)ceil_div(nb, kTile), (unsigned)ceil_div(na, kTile));
    if (ws.elem == 4) {
        int nchunks = (int)(ws.ldg / (kChunkBytes / 4));
        dist_general_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float *>(xa), na,
                                                         static_cast<const float *>(xb), nb, ws.ldg, ws.rg.ptr,
                                                         ws.ctype.ptr, nchunks, D, ldd, symmetric);
    } else {
        int nchunks = (int)(ws.ldg / (kChunkBytes / 8));
        dist_general_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double *>(xa), na,
                                                          static_cast<const double *>(xb), nb, ws.ldg, ws.rg.ptr,
                                                          ws.ctype.ptr, nchunks, D, ldd, symmetric);
    }
    FS_CUDA(cudaGetLastError());
    ++*launches;
}

}  // namespace fs
