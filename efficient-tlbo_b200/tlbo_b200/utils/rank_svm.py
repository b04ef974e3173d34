"""Pairwise RankSVM (reference tlbo/utils/rank_svm.py:7-81).

``transform_pairwise`` is the reference's double loop over ``itertools.combinations`` restated with index
arithmetic: the pairs (i < j, lexicographic = the combination order), the within-group / different-target
filter and the label balancing by the parity of the COMBINATION index k (skipped combinations count) are the
same element for element -- every output row is one exact subtraction, possibly negated.  The classifier is
scikit-learn's ``SVC(kernel='linear', C=.1)`` on the host, the very library call the reference makes."""
import numpy as np


def transform_pairwise(X, y):
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y)
    if y.ndim == 1:
        y = np.c_[y, np.ones(y.shape[0])]
    n = X.shape[0]
    i, j = np.triu_indices(n, 1)                       # combination order: k = position in this list
    k = np.arange(i.shape[0])
    keep = (y[i, 0] != y[j, 0]) & (y[i, 1] == y[j, 1])
    i, j, k = i[keep], j[keep], k[keep]
    sign = np.sign(y[i, 0] - y[j, 0])
    want = np.where(k % 2 == 0, 1.0, -1.0)             # (-1) ** k
    flip = sign != want
    Xn = X[i] - X[j]
    Xn[flip] = -Xn[flip]
    yn = np.where(flip, -sign, sign)
    return Xn, yn.ravel()


class RankSVM(object):
    def __init__(self):
        from sklearn import svm
        self.clf = svm.SVC(kernel='linear', C=.1)
        self.coef = None

    def fit(self, X, y):
        X_trans, y_trans = transform_pairwise(X, y)
        self.clf.fit(X_trans, y_trans)
        self.coef = self.clf.coef_.ravel() / np.linalg.norm(self.clf.coef_)

    def predict(self, X):
        if self.coef is not None:
            return np.dot(X, self.coef)
        raise ValueError("Must call fit() prior to predict()")

    def score(self, X, y):
        from scipy import stats
        tau, _ = stats.kendalltau(np.dot(X, self.coef), y)
        return tau
