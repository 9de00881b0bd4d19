"""Training-side neighbour lists -- restates the neighbour part of one_thread_generate_neg
(higashi/Higashi_wrapper.py:261-333), sum_duplicates (:224-233) and to_neighs_to_mask (:171-221).

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import numpy as np


def csr_row(m, row):
    """get_csr_submatrix(M,N,indptr,indices,data,row,row+1,0,N): one CSR row (cols ascending as stored)."""
    lo, hi = m.indptr[row], m.indptr[row + 1]
    return m.indices[lo:hi], m.data[lo:hi]


def sum_duplicates(col, data):
    """wrapper:224-233: sort by column, add values of equal columns."""
    order = np.argsort(col)
    col, data = col[order], data[order]
    first = np.append(True, col[1:] != col[:-1])
    return col[first], np.add.reduceat(data, np.nonzero(first)[0])


def token_neighbors(x, x_chrom, matrices, num_list, training, remove_draw,
                    nbr_cells=None, nbr_weights=None):
    """For every (cell, bin) token of every hyperedge: neighbour bins (global ids) and log1p weights.

    x int [B,3]; x_chrom int [B]; matrices(chrom, cell_1based) -> scipy CSR;
    remove_draw float [2B]: the uniform draws of wrapper:270 (token order 2r, 2r+1);
    nbr_cells/nbr_weights: per-cell kNN lists (index = 1-based cell id) or None (wrapper:281-299).
    Returns a list of 2B entries, each (ids int64, weights f32) or () when empty."""
    out = []
    B = len(x)
    for r in range(B):
        c = int(x_chrom[r])
        cell = int(x[r, 0])
        for t in range(2):
            bin_id = int(x[r, 1 + t])
            other = int(x[r, 2 - t])                     # remove_bin_ids = x[:,1:][:, ::-1]  (wrapper:264-265)
            row = bin_id - 1 - int(num_list[c])
            if nbr_cells is None:
                nbrs, val = csr_row(matrices(c, cell), row)
            else:
                cols, vals = [], []
                for nb, w in zip(nbr_cells[cell], nbr_weights[cell]):
                    cc, vv = csr_row(matrices(c, int(nb)), row)
                    cols.append(cc)
                    vals.append(vv * w)
                nbrs, val = sum_duplicates(np.concatenate(cols), np.concatenate(vals))
            if training and remove_draw[2 * r + t] >= 0.6 and len(nbrs) > 1:     # wrapper:301-305
                keep = nbrs != (other - 1 - int(num_list[c]))
                nbrs, val = nbrs[keep], val[keep]
            val = np.log1p(val)
            ids = nbrs.reshape(-1).astype(np.int64) + 1 + int(num_list[c])
            out.append((ids, val.astype(np.float32)) if len(ids) else ())
    return out


def to_neighs_to_mask(to_neighs):
    """wrapper:171-221 (COO branch): per-row L1 normalisation + relabel to first-seen unique ids.

    Returns (indices int64 [2,nnz], values f32 [nnz], unique_nodes int64 [U])."""
    uniq, uniq_list = {}, []
    rows, cols, vals = [], [], []
    for i, entry in enumerate(to_neighs):
        if len(entry) == 0:
            continue
        ids, w = entry
        w = w / np.sum(w)
        for n in ids:
            n = int(n)
            if n not in uniq:
                uniq[n] = len(uniq_list)
                uniq_list.append(n)
            cols.append(uniq[n])
            rows.append(i)
        vals.append(w)
    vals = np.concatenate(vals) if vals else np.zeros(0, np.float32)
    return (np.asarray([rows, cols], dtype=np.int64).reshape(2, -1), vals.astype(np.float32),
            np.asarray(uniq_list, dtype=np.int64))
