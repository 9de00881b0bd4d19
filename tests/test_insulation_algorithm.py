"""CPU check of the ALGORITHM behind hsg_insulation (csrc/insulation.cu), restated in numpy: fixed-point row sums, inverse-sqrt
table, the `b - a > 2w` early exit, two range updates per pair on integer difference arrays, exact prefix sums, boundary calls.
It must reproduce the vectors minted from the reference's own insulation_score / call_tads / sqrt_norm (tests/golden/insulation.npz)
-- in particular the exact zeros of the last two bins that a floating-point prefix sum would miss.  (The CUDA kernels themselves
are compared with the same vectors in tests/test_gpu_insulation.py.)"""
import numpy as np

from tests._util import load

ROW_SCALE, FIX_SCALE = float(2 ** 30), float(2 ** 40)


def device_algorithm(coords, proba, n, keep, discard, w, order):
    a, b = coords[:, 0].astype(np.int64), coords[:, 1].astype(np.int64)
    ok = (keep[a] != 0) & (keep[b] != 0)
    q = np.where(ok, np.rint(proba.astype(np.float64) * ROW_SCALE), 0.0).astype(np.int64)
    rows = np.zeros(n, dtype=np.int64)
    np.add.at(rows, a, q)
    np.add.at(rows, b, q)                                  # a diagonal pair counts twice, as m + m^T does
    r = rows.astype(np.float64) / ROW_SCALE
    inv = np.where(r > 0, 1.0 / np.sqrt(np.where(r > 0, r, 1.0)), 0.0)
    lo_, hi_ = np.minimum(a, b), np.maximum(a, b)
    use = ok & (hi_ - lo_ <= 2 * w) & (proba != 0) & (inv[lo_] != 0) & (inv[hi_] != 0)
    v = proba.astype(np.float64) * inv[lo_] * inv[hi_]
    qv = np.where(use, np.rint(v * FIX_SCALE), 0.0).astype(np.int64)
    num = np.zeros(n + 1, dtype=np.int64)
    den = np.zeros(n + 1, dtype=np.int64)
    for x, y, z in zip(lo_[qv != 0], hi_[qv != 0], qv[qv != 0]):
        if x == y:
            lo, hi = max(0, x - w), min(n - 1, x + w)
            den[lo] += 2 * z; den[hi + 1] -= 2 * z
            continue
        lo, hi = max(0, y - w), min(n - 1, x + w)
        if lo <= hi:
            den[lo] += 2 * z; den[hi + 1] -= 2 * z
        if y <= n - 2:
            lo, hi = max(x + 1, y - w), min(x + w, y - 1)
            if lo <= hi:
                num[lo] += z; num[hi + 1] -= z
    sn, sd = np.cumsum(num)[:n], np.cumsum(den)[:n]
    score = np.where(sd == 0, 1.0, sn / np.where(sd == 0, 1, sd))
    score[np.asarray(discard, dtype=np.int64)] = 1.0
    idx = np.arange(n)
    okb = (score < 0.999) & (score > 0)
    for k in range(1, order + 1):
        okb &= (score < score[np.clip(idx + k, 0, n - 1)]) & (score < score[np.clip(idx - k, 0, n - 1)])
    return score, okb.astype(np.uint8)


def test_fixed_point_range_update_algorithm_matches_reference_vectors():
    g = load("insulation.npz")
    for tag in ("a", "b", "c"):
        n, w_ins, w_tad = [int(v) for v in g["%s/cfg" % tag]]
        for c in range(g["%s/proba" % tag].shape[0]):
            score, border = device_algorithm(g["%s/coords" % tag], g["%s/proba" % tag][c], n, g["%s/keep" % tag],
                                             g["%s/discard" % tag], w_ins, w_tad)
            np.testing.assert_allclose(score, g["%s/score" % tag][c], atol=1e-7, rtol=0)
            assert np.array_equal(border, g["%s/border" % tag][c])
            assert score[n - 1] in (0.0, 1.0) and score[n - 2] in (0.0, 1.0)          # exact, not 1e-17


def test_bulk_profile_matches_reference_functions_on_upper_triangular_bulk():
    """scTAD.gen_tad's pseudo-bulk (scTAD.py:113-119): the summed vector goes into the UPPER triangle only, then the same three
    functions.  The host mirror's prefix-sum implementation against the oracle's direct restatement of those functions."""
    from higashi_b200.scTAD import bulk_profile
    from oracle import insulation as OX
    rng = np.random.Generator(np.random.PCG64(3))
    for n, band, w, order in ((90, 30, 7, 4), (41, 40, 12, 2), (64, 9, 3, 1)):
        xs, ys = np.where(np.triu(np.ones((n, n)), 1) - np.triu(np.ones((n, n)), band + 1) > 0)
        coords = np.stack([xs, ys], axis=1)
        blk = 11
        bulk = (0.1 + 3.0 * ((xs // blk) == (ys // blk)) * rng.uniform(0.5, 1.0, len(xs)) + 0.2 * rng.random(len(xs)))
        keep = np.ones(n, dtype=np.uint8); keep[rng.choice(n, 3, replace=False)] = 0
        discard = np.sort(rng.choice(n, 2, replace=False))
        score, borders = bulk_profile(coords, bulk, n, w * 10.0, order * 20.0, 10.0, keep, discard)
        m = np.zeros((n, n)); m[xs, ys] += bulk
        m *= np.outer(keep, keep)
        ref = OX.insulation_score(OX.sqrt_norm(m), w)
        ref[discard] = 1.0
        np.testing.assert_allclose(score, ref, rtol=1e-9, atol=1e-12)
        assert np.array_equal(borders, OX.call_tads(ref, order))
