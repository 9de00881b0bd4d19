"""Host-side batch producer for the training kernels (vectorised numpy).

Reference: the neighbour part of one_thread_generate_neg (Higashi_wrapper.py:261-333) + to_neighs_to_mask (:171-221):
for every (cell, bin) token its CSR row, optional 40 % removal of the partner bin while training, log1p, per-row L1
normalisation.  Output is already the CSR the kernels take (indptr over the 2B bin-token rows, chromosome-local
columns, normalised weights), so no COO / unique-node relabelling is needed."""
import numpy as np


def _rows_of(csr, cells, chrom, bins):
    rows = cells * csr.n_bins + int(csr.bin_off[chrom]) + bins
    lo, hi = csr.row_ptr[rows], csr.row_ptr[rows + 1]
    cnt = (hi - lo).astype(np.int64)
    seg = np.repeat(np.arange(len(rows)), cnt)
    idx = np.arange(int(cnt.sum()), dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt) + np.repeat(lo, cnt)
    return seg, csr.col[idx].astype(np.int64), csr.val[idx].astype(np.float32)


def neighbor_csr(csr, x, chrom, num_list, training=False, rng=None, nbr=None, nbr_w=None):
    """nbr / nbr_w: kNN cell graph ([n_cells, k+1] 1-based ids self first, weights): the token's row is the weighted sum of
    the neighbour cells' rows with duplicate columns merged (sum_duplicates, Higashi_wrapper.py:224-233, 281-293)."""
    x = np.asarray(x, dtype=np.int64)
    B = len(x)
    base = int(num_list[chrom]) + 1
    cells = np.repeat(x[:, 0] - 1, 2)
    bins = x[:, 1:3].reshape(-1) - base
    partner = x[:, [2, 1]].reshape(-1) - base
    if nbr is None:
        seg, cols, vals = _rows_of(csr, cells, chrom, bins)
    else:
        n = int(csr.bins[chrom])
        segs, colss, valss = [], [], []
        for q in range(nbr.shape[1]):
            sg, cl, vl = _rows_of(csr, nbr[cells, q] - 1, chrom, bins)
            segs.append(sg); colss.append(cl); valss.append(vl * nbr_w[cells, q][sg])
        seg, cols, vals = np.concatenate(segs), np.concatenate(colss), np.concatenate(valss)
        key = seg * n + cols
        order = np.argsort(key, kind="stable")
        key, vals = key[order], vals[order]
        first = np.append(True, key[1:] != key[:-1])
        vals = np.add.reduceat(vals, np.flatnonzero(first)).astype(np.float32) if len(key) else vals
        key = key[first]
        seg, cols = key // n, key % n
    cnt = np.bincount(seg, minlength=2 * B).astype(np.int64)
    cols = cols.astype(np.int32)
    if training:
        rng = rng or np.random.default_rng()
        drop_row = (rng.random(2 * B) >= 0.6) & (cnt > 1)
        keep = ~(drop_row[seg] & (cols == partner[seg]))
        seg, cols, vals = seg[keep], cols[keep], vals[keep]
        cnt = np.bincount(seg, minlength=2 * B)
    vals = np.log1p(vals)
    s = np.bincount(seg, weights=vals, minlength=2 * B).astype(np.float32)
    vals = vals / s[seg]
    indptr = np.zeros(2 * B + 1, dtype=np.int32)
    np.cumsum(cnt, out=indptr[1:])
    return indptr, np.ascontiguousarray(cols), np.ascontiguousarray(vals, dtype=np.float32)


def synthetic_batch(csr, chrom, num_list, n_pos, neg_num, rng, max_dist=None):
    """n_pos positives drawn from the contact matrices of one chromosome + n_pos*neg_num random negatives
    (uniform (cell, bin_i < bin_j) -- a stand-in for generate_negative_cpu, which is host code outside the kernels)."""
    n_cells, n_bins = csr.n_cells, csr.n_bins
    nc = int(csr.bins[chrom])
    base = int(num_list[chrom]) + 1
    pos = []
    while len(pos) < n_pos:
        cell = rng.integers(0, n_cells, size=n_pos)
        b1 = rng.integers(0, nc, size=n_pos)
        rows = cell * n_bins + int(csr.bin_off[chrom]) + b1
        lo, hi = csr.row_ptr[rows], csr.row_ptr[rows + 1]
        ok = hi > lo
        pick = lo + (rng.random(n_pos) * np.maximum(hi - lo, 1)).astype(np.int64)
        b2 = csr.col[np.minimum(pick, len(csr.col) - 1)]
        ok &= np.abs(b2 - b1) >= 2
        for c, i, j in zip(cell[ok], b1[ok], b2[ok]):
            pos.append((c + 1, base + min(i, j), base + max(i, j)))
    pos = np.array(pos[:n_pos], dtype=np.int64)
    n_neg = n_pos * neg_num
    cell = rng.integers(0, n_cells, size=n_neg)
    b1 = rng.integers(0, nc - 2, size=n_neg)
    b2 = np.minimum(b1 + 2 + rng.integers(0, nc, size=n_neg) % max(1, nc - 2), nc - 1)
    neg = np.stack([cell + 1, base + b1, base + b2], axis=-1).astype(np.int64)
    x = np.concatenate([pos, neg])
    y = np.concatenate([np.ones(n_pos), np.zeros(n_neg)]).astype(np.float32)
    w = np.ones(len(x), dtype=np.float32)
    perm = rng.permutation(len(x))
    return x[perm], y[perm], w[perm]
