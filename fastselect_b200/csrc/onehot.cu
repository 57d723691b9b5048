// onehot.cu -- one-hot tensor-core path (placeholder until tc_dist/tc_accum land).
#include "common.cuh"

namespace fs {

bool tensor_path_available() { return false; }

void build_onehot(fs_dataset *, WorkSet &, int *) {
    FS_REQUIRE(false, FS_ERR_STATE, "one-hot tensor path not built");
}
void launch_dist_tensor(fs_dataset *, const WorkSet &, int64_t, const int64_t *, bool, int64_t, int32_t *, int64_t,
                        cudaStream_t, int *) {
    FS_REQUIRE(false, FS_ERR_STATE, "one-hot tensor path not built");
}
void launch_accum_tensor(fs_dataset *, const WorkSet &, int, const int64_t *, const int64_t *, bool, int64_t,
                         const int8_t *, int64_t, const RowInfo *, const int32_t *, const double *, const int32_t *,
                         int32_t, double *, cudaStream_t, int *) {
    FS_REQUIRE(false, FS_ERR_STATE, "one-hot tensor path not built");
}

}  // namespace fs
