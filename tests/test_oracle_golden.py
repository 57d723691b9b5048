"""CPU: pin the oracle (oracle/fs_oracle.c) to outputs of the reference itself
(tests/golden/reference_vectors.npz, made by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import ref_oracle as R

TOL = 2e-6   # relative to max|W|: float32 accumulation-order noise of the reference


def _cases(meta, algo):
    return [m for m in meta if m["algo"] == algo and m["idx"] >= 0]


def _run(m, x, y):
    P = m["params"]
    dl = P.get("discrete_limit", 10)
    if m["algo"] == "MultiSURF":
        return R.fit_multisurf(x, y, dl, P.get("use_star", False))
    if m["algo"] == "SURF":
        return R.fit_surf(x, y, dl, P.get("use_star", False), sum_mode=2)
    return R.fit_relieff(x, y, dl, P["n_neighbors"], tie_mode=0)


@pytest.mark.parametrize("algo", ["MultiSURF", "SURF", "ReliefF"])
def test_oracle_matches_reference_vectors(golden, algo):
    arrays, meta = golden
    cases = _cases(meta, algo)
    assert len(cases) >= 20
    for m in cases:
        x, y = arrays[f"X_{m['data']}"], arrays[f"y_{m['data']}"]
        ref = arrays[f"scores_{m['idx']}"]
        scores, isd = _run(m, x, y)
        assert np.array_equal(isd, arrays[f"isd_{m['idx']}"]), m
        scale = max(1.0, float(np.abs(ref).max()))
        np.testing.assert_allclose(scores, ref, rtol=1e-5, atol=TOL * scale, err_msg=str(m))
        top = R.rank(scores, len(arrays[f"top_{m['idx']}"]))
        # identical ranking unless two adjacent reference weights are inside the noise
        ref_top = arrays[f"top_{m['idx']}"]
        if not np.array_equal(top, ref_top):
            gaps = np.abs(np.diff(np.sort(ref)[::-1][: len(ref_top) + 1]))
            assert gaps.min() < TOL * scale, m


def test_surf_sum_order_is_pinned_by_a_discriminating_case(golden):
    """On gauss2_long only the reference's real summation order (sum_mode 2) gives the
    reference's scores; the sequential and the correctly-rounded sums flip neighbour pairs."""
    arrays, meta = golden
    m = next(m for m in meta if m["data"] == "gauss2_long" and m["algo"] == "SURF" and not m["params"]["use_star"])
    x, y, ref = arrays["X_gauss2_long"], arrays["y_gauss2_long"], arrays[f"scores_{m['idx']}"]
    err = {sm: float(np.abs(R.fit_surf(x, y, 10, False, sm)[0] - ref).max()) for sm in (0, 1, 2)}
    assert err[2] == 0.0 and err[0] > 1e-5 and err[1] > 1e-5, err


def test_known_answer_vectors_of_the_survey(golden):
    """vectors captured from the reference CPU path on the
    reference's own test fixtures (tests/test_multisurf.py:19-33, tests/test_surf.py:22-32)."""
    arrays, _ = golden
    xa, ya, xb, yb = arrays["X_A"], arrays["y_A"], arrays["X_B"], arrays["y_B"]
    ka = [
        (R.fit_multisurf(xa, ya, 4, False)[0], [0.28377658, -0.35625002, -0.125, 0.0]),
        (R.fit_multisurf(xa, ya, 4, True)[0], [-3.04884768, -2.11875010, -3.40833354, 0.0]),
        (R.fit_multisurf(xa, ya, 10, False)[0], [0.2, -0.3, 0.1, 0.0]),
        (R.fit_surf(xb, yb, 3, False, 2)[0], [0.95718652, -2.0, 1.0, 0.0]),
        (R.fit_surf(xb, yb, 3, True, 2)[0], [-1.00611627, -4.0, -1.0, 0.0]),
        (R.fit_surf(xb, yb, 10, False, 2)[0], [-1, -2, 1, 0]),
        (R.fit_surf(xb, yb, 10, True, 2)[0], [-3, -4, -1, 0]),
        (R.fit_relieff(xb, yb, 4, 1)[0], [0.97247714, -1.0, 1.0, 0.0]),
        (R.fit_relieff(xb, yb, 10, 2)[0], [0, -0.5, 1, 0]),
        (R.fit_relieff(xa, ya, 4, 3)[0], [0.75709218, -0.23333332, 0.33333334, 0.0]),
    ]
    for got, want in ka:
        np.testing.assert_allclose(got, np.array(want, np.float32), rtol=1e-6, atol=1e-7)


def test_readme_quickstart(golden):
    """README.md:79-89 (config C1): MultiSURF top-15 on make_classification 500 x 1000."""
    from sklearn.datasets import make_classification

    arrays, _ = golden
    x, y = make_classification(n_samples=500, n_features=1000, n_informative=20, n_redundant=100, random_state=42)
    if not np.allclose([x.sum(), np.abs(x).sum(), float(y.sum())], arrays["readme_xsum"]):
        pytest.skip("make_classification stream differs from the one the vectors were made with")
    scores, _ = R.fit_multisurf(x, y)
    ref = arrays["readme_scores"]
    np.testing.assert_allclose(scores, ref, rtol=1e-5, atol=TOL * np.abs(ref).max())
    assert np.array_equal(R.rank(scores, 15), arrays["readme_top"])


def test_numba_argsort_restatement_orders_ties_like_numba():
    numba = pytest.importorskip("numba")

    @numba.njit
    def nb_argsort(a):
        return np.argsort(a)

    rs = np.random.RandomState(0)
    for n in (1, 2, 14, 15, 16, 17, 100, 1000, 5000):
        for levels in (2, 7, 1000):
            a = rs.randint(0, levels, n).astype(np.float32)
            a[rs.randint(0, n)] = np.inf
            assert np.array_equal(R.argsort_numba(a), nb_argsort(a)), (n, levels)


def test_threshold_roundings_decide_integer_ties(golden):
    """Fixture A with every column discrete: T_i is mathematically the integer 2 for one
    target; the reference's compiled roundings (x * 1/(n-1), fused subtraction) put it just
    below 2, so the four samples at distance 2 are NOT neighbours (naive rounding gives a
    threshold just above 2 and four more near misses) -- restated in the oracle and
    required of the GPU path."""
    arrays, _ = golden
    x32, recip, isd = R.multisurf_prep(arrays["X_A"], 10)
    out = R.multisurf_targets(x32, arrays["y_A"].astype(np.int64), recip, isd, False, np.arange(10))
    assert out["thresh"][4] < 2.0 and 2.0 - out["thresh"][4] < 1e-14
    assert out["counts"][4].tolist() == [0, 1, 0]


def test_per_target_view_sums_to_the_full_fit():
    from datasets import mixed

    x, y = mixed(11, 90, 24)
    x32, recip, isd = R.multisurf_prep(x, 10)
    yc = np.unique(y, return_inverse=True)[1].astype(np.int64)
    full = R.multisurf_scores(x32, yc, recip, isd, True)
    part = R.multisurf_targets(x32, yc, recip, isd, True, np.arange(90))["wsum"] / 90
    np.testing.assert_allclose(part, full, rtol=1e-5, atol=2e-6 * np.abs(full).max())
    x64, recip, isd = R.surf_prep(x, 10)
    full = R.surf_scores(x64, y.astype(np.int32), recip, isd, True, 2)
    part = R.surf_targets(x64, y.astype(np.int32), recip, isd, True, np.arange(90), sum_mode=2)["wsum"] / 90
    np.testing.assert_allclose(part, full, rtol=1e-5, atol=2e-6 * np.abs(full).max())
    x32, y_enc, cp, recip, isd = R.relieff_prep(x, y, 10)
    full = R.relieff_scores(x32, y_enc, recip, isd, 5, cp, 0)
    part = R.relieff_targets(x32, y_enc, recip, isd, 5, cp, np.arange(90), tie_mode=0)["wsum"] / 90
    np.testing.assert_allclose(part, full, rtol=1e-5, atol=2e-6 * np.abs(full).max())


@pytest.mark.parametrize("use_star", [False, True])
def test_byte_genotype_restatement_equals_the_float32_one(use_star):
    """fso_multisurf_targets_u8 (used at the full 20 000 x 500 000 int8 shape, where a float32 copy of the
    matrix does not fit) is the same computation as the pinned float32 restatement on all-discrete
    genotype data: identical distances, thresholds and neighbour codes, weights to float64 rounding;
    also on a column subset (TuRF's X[:, active])."""
    from datasets import epistatic_genotypes

    x, y = epistatic_genotypes(11, 300, 157)
    x[:, 5] = 1                                   # a constant column: every term 0
    tg = np.arange(0, 300, 7)
    isd = np.ones(157, bool)
    recip = np.ones(157, np.float32)
    a = R.multisurf_targets(x.astype(np.float32), y, recip, isd, use_star, tg)
    b = R.multisurf_targets_bytes(x, y, use_star, tg)
    assert np.array_equal(a["dist"], b["dist"]) and np.array_equal(a["thresh"], b["thresh"])
    assert np.array_equal(a["mask"], b["mask"])
    np.testing.assert_allclose(b["wsum"], a["wsum"], rtol=1e-13, atol=1e-13)
    cols = np.sort(np.random.RandomState(3).choice(157, 90, replace=False))
    a = R.multisurf_targets(np.ascontiguousarray(x[:, cols], np.float32), y, recip[:90], isd[:90], use_star, tg)
    b = R.multisurf_targets_bytes(x, y, use_star, tg, cols=cols)
    assert np.array_equal(a["dist"], b["dist"]) and np.array_equal(a["mask"], b["mask"])
    np.testing.assert_allclose(b["wsum"], a["wsum"], rtol=1e-13, atol=1e-13)
