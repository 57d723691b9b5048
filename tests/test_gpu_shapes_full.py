"""GPU: the reference's full benchmark shapes, where the oracle can only afford
a handful of targets: size-independent properties of the per-target outputs for hundreds of targets
(tests/properties.py -- distances symmetric / integral / zero on the diagonal, thresholds and neighbour
codes re-derived from the distances, and the feature-sum of the accumulated weights re-derived from the
distances, which ties the accumulation GEMM to the distance GEMM), additivity over row ranges, bitwise
repeatability, the planted signal, and the oracle itself on 16 targets at full width."""
import numpy as np
import pytest

from datasets import epistatic_genotypes
from oracle import ref_oracle as R
from properties import check_distance_rows, check_multisurf_rows, check_relieff_rows

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def atol_for(ref):
    return 1e-7 * max(1.0, float(np.abs(ref).max()))


def run_c3(native, n, p, n_targets, check_signal=True):
    x, y = epistatic_genotypes(42, n, p)
    isd = np.ones(p, bool)
    recip = np.full(p, 0.5, np.float32)                      # 1 / range; not used by discrete columns
    rs = np.random.RandomState(0)
    tg = np.sort(rs.choice(n, n_targets, replace=False))
    sub = tg[::16]
    split = n // 2 // 4 * 4
    with native.Dataset(x, y.astype(np.int32), 2) as ds:     # int8 end to end
        ds.set_features(isd, recip, native.FS_ARITH_F32)
        full = ds.score(native.FS_MULTISURF)
        assert np.array_equal(full, ds.score(native.FS_MULTISURF))                       # bitwise repeatable
        parts = ds.score(native.FS_MULTISURF, row_begin=0, row_end=split) + \
            ds.score(native.FS_MULTISURF, row_begin=split, row_end=n)
        np.testing.assert_allclose(parts, full, rtol=1e-12, atol=1e-9)                   # row shards add up
        got = ds.debug_rows(native.FS_MULTISURF, tg)
        got16 = ds.debug_rows(native.FS_MULTISURF, sub)
    if check_signal:
        assert set(np.argsort(full)[::-1][:2].tolist()) == {25, 75}
    check_distance_rows(got, tg, p, integer=True)
    check_multisurf_rows(got, y, tg, use_star=False)
    want = R.multisurf_targets(x.astype(np.float32), y.astype(np.int64), recip, isd, False, sub)
    assert np.array_equal(got16["dist"], want["dist"])
    assert np.array_equal(got16["thresh"], want["thresh"])
    assert np.array_equal(got16["mask"], want["mask"])
    np.testing.assert_allclose(got16["wsum"], want["wsum"], rtol=RTOL, atol=atol_for(want["wsum"]) * len(sub))
    assert np.array_equal(got["dist"][::16], got16["dist"]) and np.array_equal(got["mask"][::16], got16["mask"])


def test_c3_full_shape_multisurf_genotypes(native):
    """C3: MultiSURF on 4000 x 100 000 int8 genotypes with the epistatic label (features 25 and 75)."""
    run_c3(native, 4000, 100_000, 256)


def make_c4(n, p):
    """half 0/1/2 genotype columns, half float32 Gaussians, the epistatic label of C3
    plus a shifted continuous column -- generated in column blocks to keep the temporaries small."""
    rs = np.random.RandomState(43)
    h = p // 2
    x = np.empty((n, p), np.float32)
    x[:, :h] = rs.randint(0, 3, (n, h), dtype=np.int8)
    for c0 in range(h, p, 2500):
        c1 = min(p, c0 + 2500)
        x[:, c0:c1] = rs.standard_normal((n, c1 - c0))
    y = np.zeros(n, np.int64)
    y[(x[:, 25 % h] == 1) & (x[:, 75 % h] == 1)] = 1
    need = n // 2 - int(y.sum())
    if need > 0:
        y[rs.choice(np.flatnonzero(y == 0), need, replace=False)] = 1
    x[:, h] += 1.0 * y
    return x, y


def run_c4(native, n, p, n_targets):
    x, y = make_c4(n, p)
    h = p // 2
    # MultiSURF.fit's preprocessing (MultiSURF.py:409-420) without the per-column sort: ranges from the
    # float32 matrix, zero -> 1; the genotype half is discrete (3 values), the Gaussian half is not
    ranges = (x.max(axis=0) - x.min(axis=0)).astype(np.float32)
    ranges[ranges == 0] = 1
    recip = (1.0 / ranges).astype(np.float32)
    isd = np.arange(p) < h
    rs = np.random.RandomState(1)
    tg = np.sort(rs.choice(n, n_targets, replace=False))
    sub = tg[::8]
    with native.Dataset(x, y.astype(np.int32), 2) as ds:
        ds.set_features(isd, recip, native.FS_ARITH_F32)
        full = ds.score(native.FS_MULTISURF, use_star=True)
        got = ds.debug_rows(native.FS_MULTISURF, tg, use_star=True)
        got8 = ds.debug_rows(native.FS_MULTISURF, sub, use_star=True)
    assert np.isfinite(full).all() and full.shape == (p,)
    check_distance_rows(got, tg, p, integer=False)
    check_multisurf_rows(got, y, tg, use_star=True, tol=5e-6)
    want = R.multisurf_targets(x, y, recip, isd, True, sub)
    # float64 sums of 50 000 terms in two different orders; the threshold amplifies the difference by
    # mean^2 / variance (~1e5 at this width) through the cancellation in sum d^2 / (n - 1) - mean^2
    np.testing.assert_allclose(got8["dist"], want["dist"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got8["thresh"], want["thresh"], rtol=1e-9)
    assert np.array_equal(got8["mask"], want["mask"])
    np.testing.assert_allclose(got8["wsum"], want["wsum"], rtol=RTOL, atol=atol_for(want["wsum"]) * len(sub))


def run_c2(native, n, p, n_targets, k=10):
    rs = np.random.RandomState(7)
    y = rs.randint(0, 2, n)
    x = rs.standard_normal((n, p)).astype(np.float32)
    x[:, 0] += 1.5 * y
    x[:, 2] -= 1.0 * y
    # ReliefF.fit's preprocessing (ReliefF.py:366-380) for all-continuous columns
    x64max, x64min = x.max(axis=0).astype(np.float64), x.min(axis=0).astype(np.float64)
    recip = (1.0 / (x64max - x64min)).astype(np.float32)
    isd = np.zeros(p, bool)
    y_enc = y.astype(np.int32)
    cp = (np.bincount(y_enc) / n).astype(np.float32)
    tg = np.sort(rs.choice(n, n_targets, replace=False))
    sub = tg[::4]
    with native.Dataset(x, y_enc, 2) as ds:
        ds.set_features(isd, recip, native.FS_ARITH_F32)
        full = ds.score(native.FS_RELIEFF, k=k, class_probs=cp)
        got = ds.debug_rows(native.FS_RELIEFF, tg, k=k, class_probs=cp)
        got16 = ds.debug_rows(native.FS_RELIEFF, sub, k=k, class_probs=cp)
    assert set(np.argsort(full)[::-1][:2].tolist()) == {0, 2}
    check_distance_rows(got, tg, p, integer=False)
    check_relieff_rows(got, y_enc, cp, tg, k)
    want = R.relieff_targets(x, y_enc, recip, isd, k, cp, sub, tie_mode=0)
    # float32-rounded float64 sums of 10 000 terms: equal unless a sum sits on a rounding boundary
    np.testing.assert_allclose(got16["dist"], want["dist"], rtol=2e-7, atol=0)
    assert np.array_equal(got16["mask"], want["mask"])
    np.testing.assert_allclose(got16["wsum"], want["wsum"], rtol=RTOL, atol=atol_for(want["wsum"]) * len(sub))


def test_c2_full_shape_relieff_continuous(native):
    """C2's shape: ReliefF (k = 10) on 10 000 x 10 000 continuous columns (seeded Gaussians with two
    shifted columns instead of make_classification: the property is shape-, not data-specific)."""
    run_c2(native, 10_000, 10_000, 64)


def test_c4_full_shape_multisurf_star_mixed(native):
    """C4: MultiSURF* on 20 000 x 50 000 mixed columns (25 000 genotype + 25 000 float32 Gaussian)."""
    run_c4(native, 20_000, 50_000, 64)
