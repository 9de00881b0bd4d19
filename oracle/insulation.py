"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the single-cell insulation score / TAD boundary path that consumes the
imputed scores.  Pinned against tests/golden/insulation.npz, which was produced by executing the reference's
own function sources (tests/golden/make_golden_insulation.py).

    sqrt_norm          Higashi_analysis/Higashi_analysis.py:581-588
    insulation_score   Higashi_analysis/Higashi_TAD.py:7-19
    call_tads          Higashi_analysis/Higashi_TAD.py:22-27   (scipy argrelextrema(np.less, order, mode="clip"))
    cell_profile       scTAD.py:91-104  (matrix from coordinates + scores, mask, normalise, score, discard rows, boundaries)"""
import numpy as np


def sqrt_norm(m):
    cov = np.sqrt(m.sum(-1))
    with np.errstate(divide="ignore", invalid="ignore"):
        out = m / cov[:, None] / cov[None, :]
    out[~np.isfinite(out)] = 0.0
    return out


def insulation_score(m, w):
    """w = windowsize // res (bins).  Ratio of the cross block [i-w, i) x (i, min(n-1, i+w+1)) to the full window block."""
    n = m.shape[0]
    score = np.ones(n)
    for i in range(n):
        lo, hi = max(0, i - w), min(n, i + w + 1)
        num = m[lo:i, i + 1:min(n - 1, i + w + 1)].sum()
        den = m[lo:hi, lo:hi].sum()
        with np.errstate(divide="ignore", invalid="ignore"):
            v = num / den
        score[i] = 1.0 if np.isnan(v) else v
    return score


def call_tads(score, order):
    """order = windowsize // res // 2.  Strict local minima over +-order neighbours (edge indices clipped), 0 < score < 0.999."""
    n = len(score)
    idx = np.arange(n)
    ok = np.ones(n, dtype=bool)
    for k in range(1, order + 1):
        ok &= score < score[np.clip(idx + k, 0, n - 1)]
        ok &= score < score[np.clip(idx - k, 0, n - 1)]
    borders = np.flatnonzero(ok)
    return borders[(score[borders] < 0.999) & (score[borders] > 0)]


def cell_profile(coords, proba, n, keep, discard, w_ins, order_tad):
    m = np.zeros((n, n))
    np.add.at(m, (coords[:, 0], coords[:, 1]), proba.astype(np.float64))
    m = m + m.T
    m *= np.outer(keep, keep)
    score = insulation_score(sqrt_norm(m), w_ins)
    score[np.asarray(discard, dtype=np.int64)] = 1.0
    border = np.zeros(n, dtype=np.uint8)
    border[call_tads(score, order_tad)] = 1
    return score, border
