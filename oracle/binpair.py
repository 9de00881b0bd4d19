"""Bin-pair enumeration -- restates Impute.generate_binpair (reference higashi/Impute.py:49-61,
identical copy at higashi/Higashi_backend/utils.py:195-207) and the centromere skip set
(Impute.py:29-43, 204-208).  Integer work: results must be bit-exact.

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import numpy as np


def generate_binpair_loops(start, end, min_bin_, max_bin_, not_use_set=None):
    """Literal double loop (small cases only).  Returns int64 [P,2] of 1-based global ids."""
    if not_use_set is None:
        not_use_set = set()
    rows = []
    for bin1 in range(start, end):
        if bin1 in not_use_set:
            continue
        for bin2 in range(bin1 + min_bin_, min(bin1 + max_bin_, end)):
            if bin2 in not_use_set:
                continue
            rows.append((bin1, bin2))
    if not rows:
        return np.zeros((0, 2), dtype=np.int64)
    return np.asarray(rows, dtype=np.int64) + 1


def generate_binpair(start, end, min_bin_, max_bin_, not_use_set=None):
    """Same enumeration, vectorised with numpy (row-major, bin1 ascending then bin2 ascending).

    The band limit `bin2 < bin1 + max_bin_` is evaluated in ORIGINAL bin coordinates, and
    skipped bins are simply dropped from both positions (Impute.py:53-58).
    """
    n = int(end - start)
    if n <= 0:
        return np.zeros((0, 2), dtype=np.int64)
    use = np.ones(n, dtype=bool)
    if not_use_set:
        for b in not_use_set:
            b = int(b) - int(start)
            if 0 <= b < n:
                use[b] = False
    min_bin_ = int(min_bin_)
    max_bin_ = int(min(max_bin_, n + 1))
    b1 = np.arange(n, dtype=np.int64)
    lo = b1 + min_bin_
    hi = np.minimum(b1 + max_bin_, n)
    cnt = np.maximum(hi - lo, 0)
    cnt[~use] = 0
    tot = int(cnt.sum())
    if tot == 0:
        return np.zeros((0, 2), dtype=np.int64)
    r = np.repeat(b1, cnt)
    first = np.repeat(np.cumsum(cnt) - cnt, cnt)
    c = np.arange(tot, dtype=np.int64) - first + np.repeat(lo, cnt)
    keep = use[c]
    out = np.stack([r[keep], c[keep]], axis=-1) + int(start) + 1
    return out.astype(np.int64)


def skip_bins(acen_intervals, res):
    """Impute.skip_start_end (Impute.py:29-43): bins within +-100 kb of each `acen` cytoband.

    `acen_intervals` = iterable of (start_bp, end_bp); returns the list of local bin ids.
    """
    out = []
    for s, e in acen_intervals:
        a = int(np.floor((s - 100000) / res))
        b = int(np.ceil((e + 100000) / res))
        out.extend(range(a, b))
    return out
