"""Size-independent properties of the scoring path's per-target outputs (the dict returned by
fastselect_b200._native.Dataset.debug_rows and by the oracle's *_targets functions: dist [nt, n],
thresh [nt], mask [nt, n], wsum [p]).  Used at the reference's full benchmark shapes, where the
oracle can only afford a few targets: every check below is O(nt * n) host work.

Mask codes: 0 none, 1 near hit, 2 near miss, 3 far miss, 4 far hit (include/fastselect_b200.h)."""
import numpy as np


def check_distance_rows(out, targets, p, integer):
    d = out["dist"]
    nt = len(targets)
    assert d.shape[0] == nt
    assert (d[np.arange(nt), targets] == 0).all()                  # d_ii = 0
    assert d.min() >= 0 and d.max() <= p                            # every per-feature term is in [0, 1]
    if integer:
        assert np.array_equal(d, np.rint(d))                        # mismatch counts
    sub = d[:, targets]                                             # symmetry on the target x target block
    if integer:
        assert np.array_equal(sub, sub.T)
    else:
        np.testing.assert_allclose(sub, sub.T, rtol=2e-7, atol=1e-12)     # ReliefF / SURF rows are float32-rounded sums


def check_multisurf_rows(out, y, targets, use_star, tol=1e-12):
    """MultiSURF.py:175-251: T_i = mean - std / 2 over j != i; near = d < T_i (strict); far misses only
    with use_star; sum_f W_i[f] = (sum_near_miss d - [star] sum_far_miss d) / nM - sum_near_hit d / nH
    (a division is skipped when its count is 0) -- the accumulation checked against the distances.
    ``tol`` is relative to the size of the two terms being subtracted (they nearly cancel): 1e-12 where
    every step is exact or float64, ~5e-6 where weights are accumulated from float32 partial sums."""
    d, thresh, mask = out["dist"], out["thresh"], out["mask"]
    n = d.shape[1]
    y = np.asarray(y)
    checksum = scale = 0.0
    for r, i in enumerate(targets):
        others = np.arange(n) != i
        row = d[r]
        mu = row[others].sum() / (n - 1)
        var = max(0.0, (row[others] ** 2).sum() / (n - 1) - mu * mu)
        np.testing.assert_allclose(thresh[r], mu - 0.5 * np.sqrt(var), rtol=1e-11)
        near = (row < thresh[r]) & others
        hit = (y == y[i]) & others
        want = np.zeros(n, np.int8)
        want[near & hit] = 1
        want[near & ~hit & others] = 2
        if use_star:
            want[~near & ~hit & others] = 3
        assert np.array_equal(mask[r], want), int(i)
        n_h, n_m = int((want == 1).sum()), int((want == 2).sum())
        miss = row[want == 2].sum() - row[want == 3].sum()
        m_term = miss / n_m if n_m else miss
        h_term = row[want == 1].sum() / n_h if n_h else 0.0
        checksum += m_term - h_term
        scale += (row[want == 2].sum() + row[want == 3].sum()) / max(n_m, 1) + abs(h_term)
    assert abs(out["wsum"].sum() - checksum) <= tol * max(scale, 1.0), (out["wsum"].sum(), checksum, scale)


def check_relieff_rows(out, y_enc, class_probs, targets, k):
    """ReliefF.py:157-216: the k nearest hits and the k nearest samples of every other class (all of a
    class smaller than k); sum_f W_i[f] = -sum_hits d / h_found + sum_c P(c) / (1 - P(y_i)) sum_misses_c d / k."""
    d, mask = out["dist"], out["mask"]
    n = d.shape[1]
    y_enc = np.asarray(y_enc)
    checksum = 0.0
    for r, i in enumerate(targets):
        row = d[r]
        others = np.arange(n) != i
        denom = 1.0 - float(class_probs[y_enc[i]])
        denom = denom if denom != 0 else 1.0
        for c in range(len(class_probs)):
            members = (y_enc == c) & others
            chosen = members & (mask[r] == (1 if c == y_enc[i] else 2))
            assert int(chosen.sum()) == min(k, int(members.sum())), (int(i), c)
            if chosen.any() and (members & ~chosen).any():
                assert row[chosen].max() <= row[members & ~chosen].min(), (int(i), c)     # they are the nearest
            if c == y_enc[i]:
                checksum -= row[chosen].sum() / max(1, int(chosen.sum()))
            else:
                checksum += float(class_probs[c]) / denom * row[chosen].sum() / k
        assert not (mask[r][~others]).any()
    # the distances are float32-rounded sums, the weights sum the exact per-feature terms
    np.testing.assert_allclose(out["wsum"].sum(), checksum, rtol=1e-5, atol=1e-6)
