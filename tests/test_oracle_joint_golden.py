"""CPU: pin the joint-count oracle (oracle/fs_oracle.c) to outputs of the
reference itself (tests/golden/joint_vectors.npz, made by tests/golden/make_golden_joint.py)."""
import json
import os

import numpy as np
import pytest

from oracle import ref_oracle as R

HERE = os.path.dirname(os.path.abspath(__file__))
MI_TOL = dict(rtol=1e-12, atol=1e-15)    # reference MI is float64 (numba fastmath only reassociates)
SU_ATOL = 1e-6                           # reference SU keeps float32 probabilities / log2 terms


@pytest.fixture(scope="module")
def joint_golden():
    arrays = np.load(os.path.join(HERE, "golden", "joint_vectors.npz"))
    with open(os.path.join(HERE, "golden", "joint_vectors.json")) as fh:
        meta = json.load(fh)
    return arrays, meta


def test_mi_oracle_matches_reference(joint_golden):
    arrays, meta = joint_golden
    cases = [m for m in meta if m["kind"] == "mi"]
    assert len(cases) >= 4
    for m in cases:
        x, y = arrays[f"X_{m['data']}"], arrays[f"y_{m['data']}"]
        xe, ye, _ = R.mrmr_codes(x, y)
        for unit in ("bit", "nat"):
            rel, red = R.mi_matrices(xe, ye, unit)
            np.testing.assert_allclose(rel, arrays[f"mi_rel_{unit}_{m['data']}"], **MI_TOL)
            np.testing.assert_allclose(red, arrays[f"mi_red_{unit}_{m['data']}"], **MI_TOL)
            assert np.array_equal(red, red.T) and not red.diagonal().any()


def test_mi_is_invariant_under_recoding(joint_golden):
    """Only the partition of the samples matters: per-column codes give the same MI as the global
    coding of mRMR.py:90-92 (unused states contribute nothing, mutual_information.py:44)."""
    arrays, _ = joint_golden
    x, y = arrays["X_states"], arrays["y_states"]
    xe, ye, _ = R.mrmr_codes(x, y)
    per_col = np.stack([np.unique(x[:, f], return_inverse=True)[1] for f in range(x.shape[1])], axis=1)
    a = R.mi_matrices(xe, ye)
    b = R.mi_matrices(per_col, np.unique(y, return_inverse=True)[1])
    np.testing.assert_allclose(a[0], b[0], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-12, atol=1e-15)
    # x[:, 7] is an injective recoding of x[:, 5]: identical rows
    np.testing.assert_allclose(np.delete(a[1][5], [5, 7]), np.delete(a[1][7], [5, 7]), rtol=1e-12, atol=1e-15)


def test_su_oracle_matches_reference(joint_golden):
    arrays, meta = joint_golden
    cases = [m for m in meta if m["kind"] == "cfs"]
    assert len(cases) >= 4
    for m in cases:
        codes = arrays[f"cfs_codes_{m['data']}"]
        y_enc = np.unique(arrays[f"y_{m['data']}"], return_inverse=True)[1]
        r_cf, r_ff = R.su_matrices(codes, y_enc)
        ref_cf, ref_ff = arrays[f"cfs_rcf_{m['data']}"], arrays[f"cfs_rff_{m['data']}"]
        assert ref_cf.dtype == np.float32 and ref_ff.dtype == np.float32
        np.testing.assert_allclose(r_cf, ref_cf, rtol=0, atol=SU_ATOL)
        np.testing.assert_allclose(r_ff, ref_ff, rtol=0, atol=SU_ATOL)


def test_joint_counts_known_answers():
    a = np.array([0, 0, 1, 2, 2, 2], np.int32)
    b = np.array([1, 1, 0, 0, 1, 0], np.int32)
    t = R.joint_counts(a, b)
    assert t.tolist() == [[0, 2], [1, 0], [2, 1]]
    # independent uniform pair: MI 0; identical pair: MI = entropy = log2(3) bits, SU = 1
    x = np.array([[0, 0], [0, 1], [1, 0], [1, 1]] * 3, np.int32)
    rel, red = R.mi_matrices(x, x[:, 0])
    assert abs(red[0, 1]) < 1e-11 and abs(rel[0] - 1.0) < 1e-11
    z = np.tile(np.arange(3, dtype=np.int32), 5)[:, None]
    rel, _ = R.mi_matrices(z, z[:, 0], want_matrix=False)
    assert abs(rel[0] - np.log2(3)) < 1e-10
    r_cf, _ = R.su_matrices(z, z[:, 0], want_matrix=False)
    assert abs(r_cf[0] - 1.0) < 1e-12
    # constant column: SU 0 by the guard of CFS.py:73-74 when both entropies vanish; MI 0 up to the
    # 1e-12 the reference adds inside the logarithm (mutual_information.py:40,45)
    c = np.zeros((15, 1), np.int32)
    assert R.su_matrices(c, c[:, 0], want_matrix=False)[0][0] == 0.0
    assert abs(R.mi_matrices(c, z[:, 0], want_matrix=False)[0][0]) < 1e-11
