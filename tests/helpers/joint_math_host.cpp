// Host build of fastselect_b200/csrc/joint_math.cuh for the CPU test-suite: the very functions the
// finishing kernel of joint.cu runs per feature pair, driven by a plain loop instead of a CUDA grid.
// Input is what the device pipeline would hand them: the negated reduced-count matrix of the
// At * At^T GEMM, the reduced-row offsets and the marginal counts.  (Test helper; not part of the library.)
#include "../../fastselect_b200/csrc/joint_math.cuh"

extern "C" {

// out[pt * pt] (row-major, zero diagonal) from negC[K * ldd]
void jm_finish_host(const int32_t *negC, int64_t ldd, const int32_t *toff, int64_t pt, const int32_t *marg, int64_t n,
                    int kind, double log_base, double *out) {
    for (int64_t c = 0; c < pt; ++c) {
        out[c * pt + c] = 0.0;
        for (int64_t g = c + 1; g < pt; ++g) {
            const double v = fs::joint_pair_from_slab(negC, ldd, toff[c], toff[g], toff[c + 1] - toff[c],
                                                      toff[g + 1] - toff[g], marg + toff[c], marg + toff[g], n, kind,
                                                      log_base);
            out[c * pt + g] = v;
            out[g * pt + c] = v;
        }
    }
}

// table[16 * 16] of the pair (c, g)
void jm_table_host(const int32_t *negC, int64_t ldd, const int32_t *toff, int64_t c, int64_t g, const int32_t *marg,
                   int64_t n, int64_t *table) {
    fs::joint_table_from_slab(negC, ldd, toff[c], toff[g], toff[c + 1] - toff[c], toff[g + 1] - toff[g],
                              marg + toff[c], marg + toff[g], n, table, 16);
}

}  // extern "C"
