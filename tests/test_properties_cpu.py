"""CPU: the size-independent property checks of tests/properties.py hold for the oracle's own outputs
(so a failure of the full-size GPU tests that use them points at the kernels, not at the checks), and
they do reject corrupted outputs."""
import numpy as np
import pytest

from datasets import gaussian, genotype, mixed
from oracle import ref_oracle as R
from properties import check_distance_rows, check_multisurf_rows, check_relieff_rows


@pytest.mark.parametrize("star", [False, True])
@pytest.mark.parametrize("make,integer", [(lambda: genotype(31, 180, 90, 2), True), (lambda: mixed(32, 150, 60, 3), False)])
def test_multisurf_properties_hold_for_the_oracle(make, integer, star):
    x, y = make()
    x32, recip, isd = R.multisurf_prep(x, 10)
    yc = np.unique(y, return_inverse=True)[1].astype(np.int64)
    tg = np.sort(np.random.RandomState(0).choice(x.shape[0], 40, replace=False))
    out = R.multisurf_targets(x32, yc, recip, isd, star, tg)
    check_distance_rows(out, tg, x.shape[1], integer)
    check_multisurf_rows(out, yc, tg, star)
    bad = dict(out, wsum=out["wsum"] * 1.001)
    with pytest.raises(AssertionError):
        check_multisurf_rows(bad, yc, tg, star)
    bad = dict(out, mask=np.where(out["mask"] == 2, 1, out["mask"]).astype(np.int8))
    with pytest.raises(AssertionError):
        check_multisurf_rows(bad, yc, tg, star)


@pytest.mark.parametrize("k", [1, 10])
def test_relieff_properties_hold_for_the_oracle(k):
    x, y = gaussian(33, 200, 50, 3)
    x32, y_enc, cp, recip, isd = R.relieff_prep(x, y, 10)
    tg = np.sort(np.random.RandomState(1).choice(200, 30, replace=False))
    out = R.relieff_targets(x32, y_enc, recip, isd, k, cp, tg, tie_mode=0)
    check_distance_rows(out, tg, x.shape[1], False)
    check_relieff_rows(out, y_enc, cp, tg, k)
    bad = dict(out, wsum=out["wsum"] + 0.01)
    with pytest.raises(AssertionError):
        check_relieff_rows(bad, y_enc, cp, tg, k)


class _OracleDataset:
    """Stands in for fastselect_b200._native.Dataset so that the BODIES of the full-size GPU tests can
    be dry-run on the CPU at a small shape (test infrastructure; the GPU tests use the real library)."""

    def __init__(self, x, y_enc, n_classes):
        self.x32 = np.ascontiguousarray(x, np.float32)
        self.y = np.asarray(y_enc)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass

    def set_features(self, isd, recip, arith):
        self.isd, self.recip = np.asarray(isd, bool), np.asarray(recip, np.float32)

    def debug_rows(self, algo, targets, use_star=False, k=0, class_probs=None, feat_idx=None):
        if algo == _FakeNative.FS_MULTISURF:
            return R.multisurf_targets(self.x32, self.y.astype(np.int64), self.recip, self.isd, use_star, targets)
        return R.relieff_targets(self.x32, self.y.astype(np.int32), self.recip, self.isd, k, class_probs, targets, tie_mode=0)

    def score(self, algo, use_star=False, k=0, class_probs=None, feat_idx=None, row_begin=0, row_end=None):
        row_end = self.x32.shape[0] if row_end is None else row_end
        return self.debug_rows(algo, np.arange(row_begin, row_end), use_star, k, class_probs)["wsum"]


class _FakeNative:
    FS_RELIEFF, FS_SURF, FS_MULTISURF = 0, 1, 2
    FS_ARITH_F32, FS_ARITH_F64 = 0, 1
    Dataset = _OracleDataset


def test_full_size_test_bodies_run_clean_on_the_oracle():
    import test_gpu_shapes_full as T

    T.run_c3(_FakeNative, 320, 500, 64, check_signal=False)
    T.run_c2(_FakeNative, 300, 400, 32)
    T.run_c4(_FakeNative, 240, 300, 32)
