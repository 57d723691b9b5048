// joint_math.cuh -- per-feature-pair arithmetic of the joint-count path:
// rebuilds the full V_a x V_b contingency table of two discrete columns from the counts of their
// REDUCED one-hot rows (the tensor-core GEMM At * At^T only holds the first V - 1 values of every
// column, onehot.cu) and folds it into mutual information (mutual_information.py:35-46) or
// symmetrical uncertainty (CFS.py:26-77).
//
// Plain C++ when compiled without nvcc: tests/helpers/joint_math_host.cpp builds these same functions
// for the CPU suite, which checks them against the oracle and the reference's golden vectors.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FS_HD __host__ __device__ __forceinline__
#else
#define FS_HD inline
#endif

namespace fs {

enum : int { kJointMI = 0, kJointSU = 1 };
constexpr int kJointMaxReduced = 15;          // FS_DISTINCT_CAP - 1 reduced rows per column

// Sums needed to recover the implied last value of both columns.
struct JointSums {
    int32_t colsum[kJointMaxReduced];         // colsum[j] = sum_i c(i, j)
    int64_t sa, sb, tot;                      // sum of a's / b's reduced marginals, sum of all c(i, j)
};

// la / lb: reduced rows (V - 1) of columns a / b; ma / mb: their marginal counts;
// cnt(i, j): joint count of reduced row i of a and reduced row j of b.
template <class Cnt>
FS_HD void joint_sums(int la, int lb, const int32_t *ma, const int32_t *mb, Cnt cnt, JointSums &s) {
    s.sa = 0;
    s.sb = 0;
    s.tot = 0;
    for (int j = 0; j < lb; ++j) {
        s.colsum[j] = 0;
        s.sb += mb[j];
    }
    for (int i = 0; i < la; ++i) {
        s.sa += ma[i];
        for (int j = 0; j < lb; ++j) {
            const int32_t c = cnt(i, j);
            s.colsum[j] += c;
            s.tot += c;
        }
    }
}

// Calls visit(i, j, n_ij, n_i, n_j) for every cell of the full table in the reference's loop order
// (value codes ascending, the implied last value last: mutual_information.py:41-42, CFS.py:60-61).
//   n_ij = c(i, j)                          i < la, j < lb
//        = ma[i] - sum_j c(i, j)            j = lb   (b carries its last value)
//        = mb[j] - sum_i c(i, j)            i = la
//        = n - sa - sb + tot                both last
template <class Cnt, class Visit>
FS_HD void joint_visit(int la, int lb, const int32_t *ma, const int32_t *mb, Cnt cnt, int64_t n, const JointSums &s,
                       Visit visit) {
    for (int i = 0; i <= la; ++i) {
        const int64_t ni = i < la ? (int64_t)ma[i] : n - s.sa;
        int64_t rowsum = 0;
        for (int j = 0; j <= lb; ++j) {
            const int64_t nj = j < lb ? (int64_t)mb[j] : n - s.sb;
            int64_t nij;
            if (i < la && j < lb) {
                nij = cnt(i, j);
                rowsum += nij;
            } else if (i < la) {
                nij = ni - rowsum;
            } else if (j < lb) {
                nij = nj - s.colsum[j];
            } else {
                nij = n - s.sa - s.sb + s.tot;
            }
            visit(i, j, nij, ni, nj);
        }
    }
}

FS_HD double joint_entropy_bits(int l, const int32_t *m, int64_t n) {          // CFS.py:26-41
    int64_t rest = n;
    double e = 0.0;
    const double dn = (double)n;
    for (int i = 0; i <= l; ++i) {
        const int64_t c = i < l ? (int64_t)m[i] : rest;
        if (i < l) rest -= c;
        const double pr = (double)c / dn;
        if (pr > 1e-12) e -= pr * log2(pr);
    }
    return e;
}

// kind kJointMI: I(a; b) / log_base with the reference's 1e-12 guards (mutual_information.py:39-46);
// kind kJointSU: 2 I(a; b) / (H(a) + H(b)) in bits, 0 when both entropies vanish (CFS.py:68-77).
template <class Cnt>
FS_HD double joint_statistic(int kind, int la, int lb, const int32_t *ma, const int32_t *mb, Cnt cnt, int64_t n,
                             double log_base) {
    JointSums s;
    joint_sums(la, lb, ma, mb, cnt, s);
    // probabilities as count * (1 / n): one float64 division per cell (the ratio) instead of four -- the
    // finishing kernel is bound by float64 instructions.  count * (1 / n) and count / n differ by at most one
    // unit in the last place; the statistic moves by ~1e-16 relative (tolerance of the tests: 1e-11).
    const double inv_n = 1.0 / (double)n;
    double acc = 0.0;
    if (kind == kJointMI) {
        joint_visit(la, lb, ma, mb, cnt, n, s, [&](int, int, int64_t nij, int64_t ni, int64_t nj) {
            const double pxy = (double)nij * inv_n;
            if (pxy > 1e-12) acc += pxy * log(pxy / (((double)ni * inv_n) * ((double)nj * inv_n) + 1e-12));
        });
        return acc / log_base;
    }
    const double h = joint_entropy_bits(la, ma, n) + joint_entropy_bits(lb, mb, n);
    if (h < 1e-12) return 0.0;
    joint_visit(la, lb, ma, mb, cnt, n, s, [&](int, int, int64_t nij, int64_t ni, int64_t nj) {
        const double pxy = (double)nij * inv_n, px = (double)ni * inv_n, py = (double)nj * inv_n;
        if (pxy > 1e-12 && px > 1e-12 && py > 1e-12) acc += pxy * log2(pxy / (px * py));
    });
    return 2.0 * acc / h;
}

// One feature pair straight from the GEMM output: D holds the NEGATED reduced joint counts (the
// distance kernel of tc_dist.cu stores s_i + s_j - acc with s = 0), row ra.. of column a, column cb.. of b.
FS_HD double joint_pair_from_slab(const int32_t *D, int64_t ldd, int64_t ra, int64_t cb, int la, int lb,
                                  const int32_t *ma, const int32_t *mb, int64_t n, int kind, double log_base) {
    return joint_statistic(kind, la, lb, ma, mb, [=](int i, int j) { return -D[(ra + i) * ldd + cb + j]; }, n, log_base);
}

// Full (la + 1) x (lb + 1) table of one pair, row-major with row length ld_out (parity/debug view).
FS_HD void joint_table_from_slab(const int32_t *D, int64_t ldd, int64_t ra, int64_t cb, int la, int lb,
                                 const int32_t *ma, const int32_t *mb, int64_t n, int64_t *table, int ld_out) {
    auto cnt = [=](int i, int j) { return -D[(ra + i) * ldd + cb + j]; };
    JointSums s;
    joint_sums(la, lb, ma, mb, cnt, s);
    joint_visit(la, lb, ma, mb, cnt, n, s, [&](int i, int j, int64_t nij, int64_t, int64_t) { table[i * ld_out + j] = nij; });
}

}  // namespace fs
