"""TuRF (iterative Relief) wrapper -- host-side mirror of ``fast_select.TuRF`` (TuRF.py:7-136).

The pruning schedule, attributes and messages are the reference's.  When the base
estimator is one of this package's GPU estimators the data set is uploaded and typed
ONCE and every iteration re-scores a column subset of the resident matrix
(``fs_score(feat_idx=...)``) instead of copying ``X[:, active]`` and re-running ``fit``
(TuRF.py:110-111): per-column ranges and discreteness do not depend on which other
columns are active, so the scores are the same.
"""
from __future__ import annotations

import numpy as np
from sklearn.base import BaseEstimator, TransformerMixin, clone
from sklearn.utils.validation import check_is_fitted, validate_data

from ._relief import _narrow_integers, _ReliefBase, _validate_n_select


class TuRF(TransformerMixin, BaseEstimator):
    def __init__(self, estimator, n_features_to_select: int = 10, pct_remove: float = 0.1,
                 n_iterations: int | None = None, verbose: bool = False):
        self.estimator = estimator
        self.n_features_to_select = n_features_to_select
        self.pct_remove = pct_remove
        self.n_iterations = n_iterations
        self.verbose = verbose

    def _n_to_remove(self, n_active):
        """TuRF.py:99-102."""
        n_to_remove = max(1, int(n_active * self.pct_remove))
        if n_active - n_to_remove < self.n_features_to_select:
            n_to_remove = n_active - self.n_features_to_select
        return n_to_remove

    @staticmethod
    def _worst(scores, n_to_remove):
        """The set ``np.argsort(scores)[:n_to_remove]`` (TuRF.py:104) without a full sort: an
        O(p) selection gives the same set unless equal scores straddle the cut, in which
        case the reference's own argsort decides which of them go."""
        if n_to_remove >= scores.size:
            return np.arange(scores.size)
        kth = np.partition(scores, n_to_remove - 1)[n_to_remove - 1]
        below = scores <= kth
        if int(np.count_nonzero(below)) == n_to_remove:
            return np.flatnonzero(below)
        return np.argsort(scores)[:n_to_remove]

    def _prune(self, current_scores, score_active):
        """The pruning loop of TuRF.py:90-119.  ``score_active(active)`` returns the base
        estimator's importances for the column subset ``active``."""
        active = np.arange(self.n_features_in_)
        self.feature_importances_ = current_scores.copy()                                       # TuRF.py:88
        iteration = 0
        while True:
            if len(active) <= self.n_features_to_select:
                break
            if self.n_iterations is not None and iteration >= self.n_iterations:
                break
            n_to_remove = self._n_to_remove(len(active))
            worst = self._worst(current_scores, n_to_remove)                                    # TuRF.py:104
            active = np.delete(active, worst)
            if self.verbose:
                print(f"Iteration {iteration}: {len(active)} features remaining.")
            current_scores = score_active(active)
            iteration += 1
        order = np.argsort(current_scores)[::-1]
        self.top_features_ = np.sort(active[order])                                             # TuRF.py:117-119
        self.n_iterations_run_ = iteration
        return self

    def fit(self, X, y):
        resident = isinstance(self.estimator, _ReliefBase)
        if resident:
            # keep int8/uint8/float32 matrices as they are: the base estimator's own
            # validation decides the arithmetic, exactly as a direct fit would
            Xv, y = validate_data(self, _narrow_integers(X), y, y_numeric=True,
                                  dtype=[np.float64, np.float32, np.int8, np.uint8], ensure_2d=True)
        else:
            Xv, y = validate_data(self, X, y, y_numeric=True, dtype=np.float64, ensure_2d=True)   # TuRF.py:77-79
        self.n_features_in_ = Xv.shape[1]
        if not 0 < self.pct_remove < 1:
            raise ValueError("pct_remove must be between 0 and 1.")

        base_estimator = clone(self.estimator)
        if not resident:
            def refit(active):
                base_estimator.fit(Xv[:, active], y)                                            # TuRF.py:110-111
                return base_estimator.feature_importances_

            base_estimator.fit(Xv, y)
            return self._prune(base_estimator.feature_importances_, refit)

        session, _ = base_estimator._open_session(Xv, y)
        if session is None:         # single-class ReliefF: zero scores (ReliefF.py:351-356)
            return self._prune(base_estimator.feature_importances_,
                               lambda active: np.zeros(len(active), dtype=np.float32))
        with session:
            return self._prune(session.score(), self._checked_scorer(session, base_estimator))

    @staticmethod
    def _checked_scorer(session, base_estimator):
        """Scores a column subset of the resident data set after the check the reference's per-iteration
        ``base_estimator.fit(X[:, active], y)`` (TuRF.py:110-111) would have made: an integer
        ``n_features_to_select`` of the base estimator larger than the remaining columns is an error."""
        def score_active(active):
            _validate_n_select(base_estimator.n_features_to_select, len(active))
            return session.score(active)

        return score_active

    def transform(self, X):
        check_is_fitted(self)
        X = validate_data(self, X, reset=False, dtype=[np.float64, np.float32])
        return X[:, self.top_features_]

    def fit_transform(self, X, y=None, **fit_params):
        self.fit(X, y)
        return self.transform(X)
