// Probe (not part of the library): cost of reading a 32-lane x 16-column chunk of TMEM with different
// tcgen05.ld shapes, as the accumulation epilogue does it -- 12 warps per CTA (3 per lane quarter), one CTA per
// SM, no MMA running.  "lat": load, wait, use, repeat; "pipe": the next chunk's loads are issued before the
// current chunk is used (what the kernel does).  Also prints the register layout of 16x256b.x2.
//   make -C fastselect_b200/csrc probe4 && tools/build/tmem_ld_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "../fastselect_b200/csrc/tc_common.cuh"

using namespace fs::tc;

__device__ __forceinline__ void ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

// one chunk = 32 lanes x 16 columns = 16 registers per thread
template <int kShape>
__device__ __forceinline__ void load_chunk(uint32_t tq, uint32_t col, uint32_t (&v)[16]) {
    if constexpr (kShape == 0) {
        tmem_ld_32x16(tq + col, v);
    } else if constexpr (kShape == 1) {
        uint32_t a[4], b[4], c[4], d[4];
        tmem_ld_16x256b(tq + col, a);
        tmem_ld_16x256b(tq + col + 8, b);
        tmem_ld_16x256b(tq + (16u << 16) + col, c);
        tmem_ld_16x256b(tq + (16u << 16) + col + 8, d);
        for (int i = 0; i < 4; ++i) { v[i] = a[i]; v[4 + i] = b[i]; v[8 + i] = c[i]; v[12 + i] = d[i]; }
    } else if constexpr (kShape == 2) {
        uint32_t a[8], b[8];
        ld_16x256b_x2(tq + col, a);
        ld_16x256b_x2(tq + (16u << 16) + col, b);
        for (int i = 0; i < 8; ++i) { v[i] = a[i]; v[8 + i] = b[i]; }
    } else if constexpr (kShape == 3) {
        uint32_t a[8], b[8];
        ld_32x32b_x8(tq + col, a);
        ld_32x32b_x8(tq + col + 8, b);
        for (int i = 0; i < 8; ++i) { v[i] = a[i]; v[8 + i] = b[i]; }
    } else {
        // 16x256b.x4 covers 16 lanes x 32 columns: half of the lanes of TWO chunks; callers use it pairwise
        ld_16x256b_x4(tq + col, v);
    }
}

template <int kShape, bool kPipe>
__global__ void __launch_bounds__(448, 1) bench(long long *cycles, uint32_t *sink, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 1) tmem_alloc<512>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot;
    if (warp >= 2) {
        const uint32_t tq = base + ((uint32_t)((warp & 3) * 32) << 16);
        const int part = (warp - 2) >> 2;
        uint32_t acc = 0;
        uint32_t v[2][16];
        const long long t0 = clock64();
        if (kPipe) load_chunk<kShape>(tq, (uint32_t)(16 * part), v[0]);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const uint32_t col = (uint32_t)(16 * (part + 3 * i) + 240 * (it & 1));
                if (kPipe) {
                    tmem_ld_wait();
                    const uint32_t ncol = (uint32_t)(16 * (part + 3 * ((i + 1) % 5)) + 240 * ((it + (i == 4)) & 1));
                    load_chunk<kShape>(tq, ncol, v[(i + 1) & 1]);
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc += v[i & 1][e] * (uint32_t)(e + 1);
                } else {
                    load_chunk<kShape>(tq, col, v[0]);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc += v[0][e] * (uint32_t)(e + 1);
                }
            }
        }
        tmem_ld_wait();
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 12 + warp - 2] = t1 - t0;
        sink[blockIdx.x * 448 + threadIdx.x] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(base);
    }
}

__global__ void __launch_bounds__(128, 1) layout_x2(uint32_t *out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<32>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot;
    for (int c = 0; c < 32; ++c)
        tmem_st_32x1(base + ((uint32_t)(warp * 32) << 16) + c, (uint32_t)((warp * 32 + lane) * 1000 + c));
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        uint32_t r[8];
        ld_16x256b_x2(base, r);
        tmem_ld_wait();
        for (int i = 0; i < 8; ++i) out[lane * 8 + i] = r[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<32>(base);
    }
}

template <int kShape, bool kPipe>
static void run(const char *name, long long *d_cycles, uint32_t *d_sink, int sms) {
    const int iters = 2000;
    bench<kShape, kPipe><<<sms, 448>>>(d_cycles, d_sink, 10);
    bench<kShape, kPipe><<<sms, 448>>>(d_cycles, d_sink, iters);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: failed\n", name); return; }
    static long long h[148 * 12];
    cudaMemcpy(h, d_cycles, sizeof(long long) * sms * 12, cudaMemcpyDeviceToHost);
    double s = 0, mx = 0;
    for (int i = 0; i < sms * 12; ++i) { s += (double)h[i]; if ((double)h[i] > mx) mx = (double)h[i]; }
    printf("%-34s %-5s  %7.1f cycles per chunk (32 lanes x 16 columns) per warp, avg; %7.1f max; => %6.0f cycles per 240-column item\n",
           name, kPipe ? "pipe" : "lat", s / (sms * 12) / (iters * 5.0), mx / (iters * 5.0), s / (sms * 12) / iters);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (sms > 148) sms = 148;
    long long *d_cycles;
    uint32_t *d_sink, *d_lay, h_lay[256];
    cudaMalloc(&d_cycles, sizeof(long long) * 148 * 12);
    cudaMalloc(&d_sink, sizeof(uint32_t) * 148 * 448);
    cudaMalloc(&d_lay, sizeof(h_lay));
    layout_x2<<<1, 128>>>(d_lay);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("layout probe failed\n"); return 1; }
    cudaMemcpy(h_lay, d_lay, sizeof(h_lay), cudaMemcpyDeviceToHost);
    printf("16x256b.x2 at lane offset 0: thread -> (lane, column) of r0..r7\n");
    for (int t = 0; t < 32; t += 5) {
        printf("  t%02d:", t);
        for (int i = 0; i < 8; ++i) printf(" (%2u,%2u)", h_lay[t * 8 + i] / 1000, h_lay[t * 8 + i] % 1000);
        printf("\n");
    }
    run<0, false>("32x32b.x16 (1 instruction)", d_cycles, d_sink, sms);
    run<0, true>("32x32b.x16 (1 instruction)", d_cycles, d_sink, sms);
    run<3, false>("32x32b.x8 x 2", d_cycles, d_sink, sms);
    run<3, true>("32x32b.x8 x 2", d_cycles, d_sink, sms);
    run<1, false>("16x256b.x1 x 4", d_cycles, d_sink, sms);
    run<1, true>("16x256b.x1 x 4", d_cycles, d_sink, sms);
    run<2, false>("16x256b.x2 x 2", d_cycles, d_sink, sms);
    run<2, true>("16x256b.x2 x 2", d_cycles, d_sink, sms);
    run<4, false>("16x256b.x4 (half the lanes, 32 col)", d_cycles, d_sink, sms);
    run<4, true>("16x256b.x4 (half the lanes, 32 col)", d_cycles, d_sink, sms);
    return 0;
}
