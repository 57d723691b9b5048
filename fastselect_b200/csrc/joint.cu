// joint.cu -- joint-count path: pairwise contingency tables of discrete
// columns on the tensor cores, folded into mutual information or symmetrical uncertainty.
//
// Replaces _batch_mi_cpu / calculate_mi_matrices (mutual_information.py:49-63, :158-196; the reference
// leaves the p x p redundancy matrix to the CPU even with backend="gpu", :191-193) and CFS's
// _precompute_correlations_cpu / _precompute_correlations_gpu_kernel (CFS.py:81-104, :219-243: one
// thread per feature walking every sample of every other feature).
//
// Data flow (all on the data set's stream):
//   encode   the reduced one-hot rows At [K, n] of the columns -- the feature-major FP4 operand the
//            accumulation GEMM already uses (onehot.cu); a column with V values owns V - 1 rows.
//   marginals  popcount of every At row (HBM-bound, K * n / 2 bytes).
//   counts   C = At * At^T, contraction over the samples: the distance kernel of tc_dist.cu is
//            reused as it is (tcgen05 kind::mxf4, FP32 accumulators, exact below 2^24 samples) with
//            At as both operands and a zero s vector, so a band of C arrives NEGATED as int32 rows.
//            Only the upper triangle is needed: a band of rows [b0, b0 + R) is multiplied with the
//            rows b0 .. K only.  Tensor-pipe bound: 2 * K^2/2 * n operations for the whole matrix.
//   finish   one thread per column pair rebuilds the full V_a x V_b table from the reduced counts
//            and the marginals (joint_math.cuh) and writes the statistic to both triangles of the
//            p x p result.  Reads 4 * (V_a - 1)(V_b - 1) bytes per pair, writes 16.
// Bands bound the slab (FS_B200_JOINT_SLAB_MB, default 1024 MB) whatever K is.
#include <algorithm>
#include <cstring>

#include <cuda.h>

#include "common.cuh"
#include "joint_math.cuh"

namespace fs {

// onehot.cu / tc_dist.cu
CUtensorMap make_tmap_u8_sw128(const void *base, uint64_t row_bytes, uint64_t rows, uint64_t pitch_bytes,
                               uint32_t box_rows);
void launch_tc_dist(const CUtensorMap &tmap_a, const CUtensorMap &tmap_b, const CUtensorMap *tmap_b_half, int64_t K,
                    const int32_t *srow, const int64_t *d_ids, int64_t R, int64_t n, int64_t ldd, bool symmetric, bool subtract,
                    const DistPeers &peers, cudaStream_t st, int *launches, double *ops);

namespace {

// marg[k] = number of samples carrying reduced row k: its 1.0 nibbles (0x2) in At.  One warp per row.
__global__ void __launch_bounds__(256) joint_marginals_kernel(const int8_t *__restrict__ At, int64_t pitch,
                                                              int64_t row_bytes, int64_t K, int32_t *__restrict__ marg) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= K) return;
    const uint8_t *r = reinterpret_cast<const uint8_t *>(At) + row * pitch;      // pitch: multiple of 64 bytes
    const int64_t vecs = row_bytes >> 4;
    int cnt = 0;
    for (int64_t w = lane; w < vecs; w += 32) {
        const uint4 q = reinterpret_cast<const uint4 *>(r)[w];
        cnt += __popc(q.x & 0x22222222u) + __popc(q.y & 0x22222222u) + __popc(q.z & 0x22222222u) + __popc(q.w & 0x22222222u);
    }
    for (int64_t b = (vecs << 4) + lane; b < row_bytes; b += 32) cnt += __popc((unsigned)r[b] & 0x22u);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) marg[row] = cnt;
}

// out[c, g] = out[g, c] = statistic of the columns at positions c in [c_lo, c_hi) and g in (c, n_kept).
// start[c] .. start[c + 1]: reduced rows of position c (an empty range for a constant column);
// D: the band's negated counts, its row 0 / column 0 being reduced row b0.
__global__ void __launch_bounds__(256) joint_finish_kernel(const int32_t *__restrict__ D, int64_t ldd, int64_t b0,
                                                           const int32_t *__restrict__ start, int64_t c_lo, int64_t c_hi,
                                                           int64_t n_kept, const int32_t *__restrict__ marg, int64_t n,
                                                           int kind, double log_base, double *__restrict__ out) {
    const int64_t g = c_lo + 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_kept) return;
    const int sg = start[g], lg = start[g + 1] - sg;
    for (int64_t c = c_lo + blockIdx.y; c < c_hi && c < g; c += gridDim.y) {
        const int sc = start[c], lc = start[c + 1] - sc;
        const double v = joint_pair_from_slab(D, ldd, sc - b0, sg - b0, lc, lg, marg + sc, marg + sg, n, kind, log_base);
        out[c * n_kept + g] = v;
        out[g * n_kept + c] = v;
    }
}

// Parity/debug view: full tables of the requested pairs whose smaller position lies in [c_lo, c_hi).
// tables[q * 256 + va * 16 + vb] with (a, b) in the caller's order.
__global__ void __launch_bounds__(128) joint_tables_kernel(const int32_t *__restrict__ D, int64_t ldd, int64_t b0,
                                                           const int32_t *__restrict__ start, int64_t c_lo, int64_t c_hi,
                                                           const int32_t *__restrict__ marg, int64_t n,
                                                           const int64_t *__restrict__ pairs, int64_t m,
                                                           int64_t *__restrict__ tables) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const int64_t a = pairs[2 * q], b = pairs[2 * q + 1];
    const int64_t c = a < b ? a : b, g = a < b ? b : a;
    if (c < c_lo || c >= c_hi) return;
    const int sc = start[c], lc = start[c + 1] - sc, sg = start[g], lg = start[g + 1] - sg;
    const int64_t ra = sc - b0, cb = sg - b0;
    auto cnt = [=](int i, int j) { return -D[(ra + i) * ldd + cb + j]; };
    JointSums s;
    joint_sums(lc, lg, marg + sc, marg + sg, cnt, s);
    int64_t *t = tables + q * 256;
    const bool swapped = a > b;
    joint_visit(lc, lg, marg + sc, marg + sg, cnt, n, s, [&](int i, int j, int64_t nij, int64_t, int64_t) {
        t[swapped ? j * 16 + i : i * 16 + j] = nij;
    });
}

struct JointTimer {
    bool on;
    cudaStream_t st;
    struct Span { int ph; cudaEvent_t a, b; };
    std::vector<Span> spans;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    JointTimer(bool enable, cudaStream_t s) : on(enable), st(s) {
        if (on) {
            cudaEventCreate(&t0);
            cudaEventCreate(&t1);
            cudaEventRecord(t0, st);
        }
    }
    void begin(int ph) {
        if (!on) return;
        Span s{ph, nullptr, nullptr};
        cudaEventCreate(&s.a);
        cudaEventCreate(&s.b);
        cudaEventRecord(s.a, st);
        spans.push_back(s);
    }
    void end() {
        if (on) cudaEventRecord(spans.back().b, st);
    }
    void finish(float *ms_total, float *ms_phase, int n_phase) {
        if (!on) return;
        cudaEventRecord(t1, st);
        cudaEventSynchronize(t1);
        for (int i = 0; i < n_phase; ++i) ms_phase[i] = 0.f;
        for (auto &s : spans) {
            float t = 0.f;
            cudaEventElapsedTime(&t, s.a, s.b);
            ms_phase[s.ph] += t;
        }
        cudaEventElapsedTime(ms_total, t0, t1);
    }
    ~JointTimer() {
        for (auto &s : spans) {
            cudaEventDestroy(s.a);
            cudaEventDestroy(s.b);
        }
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
    }
};

// The GEMM bands shared by fs_joint_matrix (kind >= 0, d_out) and fs_joint_tables (d_pairs / d_tables).
void run_joint(fs_dataset *ds, int kind, double log_base, const int64_t *feat_idx, int64_t n_kept, int64_t pos_begin,
               int64_t pos_end, double *d_out, const int64_t *d_pairs, int64_t m, int64_t *d_tables, fs_stats *stats) {
    FS_REQUIRE(ds->have_features, FS_ERR_STATE, "joint counts: call fs_dataset_set_features first");
    FS_REQUIRE(n_kept >= 1 && pos_begin >= 0 && pos_begin <= pos_end && pos_end <= n_kept, FS_ERR_INVALID,
               "joint counts: bad position range [%lld, %lld) of %lld columns", (long long)pos_begin, (long long)pos_end,
               (long long)n_kept);
    // FP32 accumulators hold the counts exactly below 2^24 samples
    FS_REQUIRE(ds->n < (1LL << 24), FS_ERR_INVALID, "joint counts: n = %lld exceeds the exact range (2^24)", (long long)ds->n);
    const int64_t n = ds->n;
    cudaStream_t st = ds->stream;
    // reduced rows of every position: V - 1 for a tensor-path column, none for a constant one
    std::vector<int32_t> start(n_kept + 1, 0);
    for (int64_t c = 0; c < n_kept; ++c) {
        const int64_t f = feat_idx ? feat_idx[c] : c;
        FS_REQUIRE(f >= 0 && f < ds->p, FS_ERR_INVALID, "feat_idx[%lld]=%lld outside [0,%lld)", (long long)c, (long long)f,
                   (long long)ds->p);
        const unsigned ci = ds->col_info[f], path = ci & kColPathMask;
        FS_REQUIRE(path == kColTensor || path == kColConst, FS_ERR_INVALID,
                   "joint counts: column %lld is not a discrete column with at most %d distinct values", (long long)f,
                   FS_DISTINCT_CAP);
        start[c + 1] = start[c] + (path == kColTensor ? (int32_t)(ci >> 4) : 0);
    }
    int launches = 0;
    double ops = 0.0;
    JointTimer tm(stats != nullptr, st);
    enum { PH_ENCODE = 0, PH_COUNTS, PH_FINISH, PH_N };
    tm.begin(PH_ENCODE);
    ds->no_dist_ops = true;
    try {
        build_workset(ds, feat_idx, n_kept, true, false, 0, n, true, false, false, false, &launches);
    } catch (...) {
        ds->no_dist_ops = false;
        throw;
    }
    ds->no_dist_ops = false;
    const WorkSet &ws = ds->ws;
    const int64_t K = ws.K_used;
    FS_REQUIRE(K == start[n_kept], FS_ERR_STATE, "joint counts: working set holds %lld reduced rows, expected %lld",
               (long long)K, (long long)start[n_kept]);
    ds->joint_start.reserve(n_kept + 1);
    // pageable source: the copy is staged before the call returns
    FS_CUDA(cudaMemcpyAsync(ds->joint_start.ptr, start.data(), (n_kept + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    const int64_t row_bytes = (n + 1) / 2, pitch = ws.ldt / 2;     // At: two samples per byte
    ds->joint_marg.reserve(std::max<int64_t>(K, 1));
    if (K > 0) {
        joint_marginals_kernel<<<(unsigned)ceil_div(K, 8), 256, 0, st>>>(ws.At.ptr, pitch, row_bytes, K, ds->joint_marg.ptr);
        FS_CUDA(cudaGetLastError());
        ++launches;
    }
    tm.end();

    // the reused distance kernel adds s_i + s_j (zeros here) and looks its rows up through an id vector
    const int64_t zero_len = round_up(std::max<int64_t>(K, 1), 256) + 256;
    ds->joint_zero.reserve(zero_len);
    FS_CUDA(cudaMemsetAsync(ds->joint_zero.ptr, 0, zero_len * sizeof(int32_t), st));
    const char *env = getenv("FS_B200_JOINT_SLAB_MB");
    const int64_t budget = std::max<int64_t>(1, env ? atoll(env) : 1024) << 20;
    const int64_t k_bytes = round_up(row_bytes, 128);              // contraction length walked by the kernel (TMA zero-fills)
    int64_t c_lo = pos_begin;
    int bands = 0;
    while (c_lo < pos_end) {
        ++bands;
        const int64_t b0 = start[c_lo], ncols = K - b0;
        int64_t c_hi = pos_end, R = 0, ldd = 128;
        if (ncols > 0) {
            ldd = round_up(ncols, 128);
            const int64_t max_rows = std::max<int64_t>(128, budget / (4 * ldd) / 128 * 128);
            c_hi = c_lo + 1;
            while (c_hi < pos_end && start[c_hi + 1] - b0 <= max_rows) ++c_hi;
            R = start[c_hi] - b0;
        }
        if (R > 0) {
            tm.begin(PH_COUNTS);
            ds->joint_slab.reserve((size_t)round_up(R, 128) * ldd);
            ds->joint_ids.reserve(round_up(R, 128));
            FS_CUDA(cudaMemsetAsync(ds->joint_ids.ptr, 0, round_up(R, 128) * sizeof(int64_t), st));
            const int8_t *base = ws.At.ptr + (size_t)b0 * pitch;
            const CUtensorMap ta = make_tmap_u8_sw128(base, row_bytes, R, pitch, 128);
            const CUtensorMap tb = make_tmap_u8_sw128(base, row_bytes, ncols, pitch, 256);
            DistPeers peers{};
            peers.world = 1;
            peers.rank = 0;
            peers.slab[0] = ds->joint_slab.ptr;
            peers.starts[0] = 0;
            peers.starts[1] = ncols;
            peers.sb_base[0] = 0;
            peers.sb_base[1] = (int32_t)ceil_div(ncols, 256);
            peers.coarse_shift = 3;
            const CUtensorMap tbh = make_tmap_u8_sw128(base, row_bytes, ncols, pitch, 128);      // CTA pairs
            launch_tc_dist(ta, tb, &tbh, k_bytes, ds->joint_zero.ptr, ds->joint_ids.ptr, R, ncols, ldd, false, false, peers, st,
                           &launches, &ops);
            tm.end();
        }
        tm.begin(PH_FINISH);
        if (d_out && c_lo + 1 < n_kept) {
            dim3 grid((unsigned)ceil_div(n_kept - c_lo - 1, 256), (unsigned)std::min<int64_t>(c_hi - c_lo, 4096));
            joint_finish_kernel<<<grid, 256, 0, st>>>(ds->joint_slab.ptr, ldd, b0, ds->joint_start.ptr, c_lo, c_hi, n_kept,
                                                      ds->joint_marg.ptr, n, kind, log_base, d_out);
            FS_CUDA(cudaGetLastError());
            ++launches;
        }
        if (d_tables && m > 0) {
            joint_tables_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, st>>>(ds->joint_slab.ptr, ldd, b0, ds->joint_start.ptr,
                                                                            c_lo, c_hi, ds->joint_marg.ptr, n, d_pairs, m,
                                                                            d_tables);
            FS_CUDA(cudaGetLastError());
            ++launches;
        }
        tm.end();
        c_lo = c_hi;
    }
    // `start` (pageable) was the source of an asynchronous copy
    FS_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        float ms[PH_N];
        tm.finish(&stats->ms_total, ms, PH_N);
        stats->ms_gather = ms[PH_ENCODE];
        stats->ms_dist_tensor = ms[PH_COUNTS];
        stats->ms_reduce = ms[PH_FINISH];
        stats->launches = launches;
        stats->n_chunks = bands;
        stats->n_tensor_cols = ws.pt;
        stats->onehot_k = K;
        stats->ops_dist_tensor = ops;
    }
}

}  // namespace
}  // namespace fs

using namespace fs;

extern "C" {

int fs_joint_matrix(fs_dataset *ds, int kind, double log_base, const int64_t *feat_idx, int64_t n_kept,
                    int64_t pos_begin, int64_t pos_end, double *out, int out_on_device, fs_stats *stats) {
    try {
        FS_REQUIRE(ds && out, FS_ERR_INVALID, "fs_joint_matrix: null pointer");
        FS_REQUIRE(kind == FS_JOINT_MI || kind == FS_JOINT_SU, FS_ERR_INVALID, "fs_joint_matrix: unknown kind %d", kind);
        FS_REQUIRE(kind != FS_JOINT_MI || log_base > 0.0, FS_ERR_INVALID, "fs_joint_matrix: log_base must be positive");
        if (!feat_idx) n_kept = ds->p;
        FS_REQUIRE(n_kept >= 1, FS_ERR_INVALID, "fs_joint_matrix: no columns");
        FS_CUDA(cudaSetDevice(ds->device));
        alloc_stream() = ds->stream;
        double *d_out = out;
        const size_t cells = (size_t)n_kept * (size_t)n_kept;
        if (!out_on_device) {
            ds->joint_out.reserve(cells);
            d_out = ds->joint_out.ptr;
        }
        FS_CUDA(cudaMemsetAsync(d_out, 0, cells * sizeof(double), ds->stream));
        run_joint(ds, kind, log_base, feat_idx, n_kept, pos_begin, pos_end, d_out, nullptr, 0, nullptr, stats);
        if (!out_on_device) {
            FS_CUDA(cudaMemcpyAsync(out, d_out, cells * sizeof(double), cudaMemcpyDeviceToHost, ds->stream));
            FS_CUDA(cudaStreamSynchronize(ds->stream));
        }
        return FS_OK;
    } catch (const Fail &f) {
        return f.code;
    } catch (const std::exception &e) {
        set_error("fs_joint_matrix: %s", e.what());
        return FS_ERR_OOM;
    }
}

int fs_joint_tables(fs_dataset *ds, const int64_t *feat_idx, int64_t n_kept, const int64_t *pairs, int64_t m,
                    int64_t *tables_out) {
    try {
        FS_REQUIRE(ds && pairs && tables_out && m >= 1, FS_ERR_INVALID, "fs_joint_tables: invalid argument");
        if (!feat_idx) n_kept = ds->p;
        for (int64_t q = 0; q < m; ++q)
            FS_REQUIRE(pairs[2 * q] >= 0 && pairs[2 * q] < n_kept && pairs[2 * q + 1] >= 0 && pairs[2 * q + 1] < n_kept &&
                           pairs[2 * q] != pairs[2 * q + 1],
                       FS_ERR_INVALID, "fs_joint_tables: pair %lld = (%lld, %lld) is not two different positions below %lld",
                       (long long)q, (long long)pairs[2 * q], (long long)pairs[2 * q + 1], (long long)n_kept);
        FS_CUDA(cudaSetDevice(ds->device));
        alloc_stream() = ds->stream;
        ds->joint_pairs.reserve(2 * m);
        ds->joint_tables.reserve(256 * m);
        FS_CUDA(cudaMemcpyAsync(ds->joint_pairs.ptr, pairs, 2 * m * sizeof(int64_t), cudaMemcpyHostToDevice, ds->stream));
        FS_CUDA(cudaMemsetAsync(ds->joint_tables.ptr, 0, 256 * m * sizeof(int64_t), ds->stream));
        run_joint(ds, -1, 1.0, feat_idx, n_kept, 0, n_kept, nullptr, ds->joint_pairs.ptr, m, ds->joint_tables.ptr, nullptr);
        FS_CUDA(cudaMemcpyAsync(tables_out, ds->joint_tables.ptr, 256 * m * sizeof(int64_t), cudaMemcpyDeviceToHost, ds->stream));
        FS_CUDA(cudaStreamSynchronize(ds->stream));
        return FS_OK;
    } catch (const Fail &f) {
        return f.code;
    } catch (const std::exception &e) {
        set_error("fs_joint_tables: %s", e.what());
        return FS_ERR_OOM;
    }
}

}  // extern "C"
