"""Imputation driver -- restates Impute.impute_process (higashi/Impute.py:123-302): pair enumeration
per chromosome, per-cell prep_one + fix_cell2 + predict(batch 1e5), payload layout
(`coordinates` int64 [P_c,2] local 0-based; `cell_{id}` float32 [P_c]).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Also the CPU baseline timed by bench.py.
"""
import numpy as np
import torch

from . import binpair, prep, scorer


def enumerate_pairs(num_list, chroms, min_bin, max_bin, skip=None):
    """Impute.py:186-237.  Returns (big_samples int64 [P,3] with column 0 = 1, chrom id [P],
    {chrom: (slice_start, slice_end)}, {chrom: coordinates int64 [P_c,2]})."""
    big, big_chrom, slices, coords = [], [], {}, {}
    s0 = 0
    for j in chroms:
        not_use = set()
        if skip is not None:
            for (cj, a, b) in skip:
                if int(cj) == j:
                    for bin_ in range(int(a), int(b)):
                        not_use.add(bin_ + int(num_list[j]))
        sb = binpair.generate_binpair(int(num_list[j]), int(num_list[j + 1]), min_bin, max_bin, not_use)
        coords[j] = sb - int(num_list[j]) - 1
        big.append(np.concatenate([np.ones((len(sb), 1), dtype=np.int64), sb], axis=-1))
        big_chrom.append(np.full(len(sb), j, dtype=np.int64))
        slices[j] = (s0, s0 + len(sb))
        s0 += len(sb)
    return np.concatenate(big), np.concatenate(big_chrom), slices, coords


def impute_cells(sd, tables, cell_feats, num_list, matrices, cells, chroms, big_samples,
                 local_transfer_range=0, nbr=None, nbr_w=None, activation="sigmoid", batch_size=100000):
    """Hot loop of impute_process (Impute.py:253-277) for 0-based `cells`.

    matrices(chrom, cell_1based) -> scipy CSR; nbr/nbr_w: [n_cells, k+1] arrays (1-based, self first)
    or None.  Returns float32 [len(cells), P]."""
    out = np.zeros((len(cells), len(big_samples)), dtype=np.float32)
    x = big_samples.copy()
    with torch.no_grad():
        for n, i in enumerate(cells):
            cell = int(i) + 1
            adj = []
            for j in chroms:
                if nbr is None:
                    a = prep.prep_one(lambda c, j=j: matrices(j, c), local_transfer_range, [cell], None)
                else:
                    a = prep.prep_one(lambda c, j=j: matrices(j, c), local_transfer_range,
                                      list(nbr[i]), list(nbr_w[i]))
                adj.append(a)
            # fix_cell2 fills the chromosomes of impute_list; other chromosomes are never scored
            full_adj = adj if len(chroms) == len(num_list) - 1 else None
            if full_adj is None:
                raise NotImplementedError("oracle imputes all chromosomes of chrom_list")
            bin_feats = scorer.fix_cell(adj, tables, num_list)
            x[:, 0] = cell
            out[n] = scorer.predict(x, bin_feats, num_list, sd, tables, cell_feats, activation,
                                    batch_size).reshape(-1)
    return out
