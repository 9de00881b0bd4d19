"""Shared helpers for the tests: load golden fixtures, rebuild datasets from them."""
import os

import numpy as np

from higashi_b200.synthetic import PackedCSR

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


class FixtureDataset:
    """The `ds/*` arrays of a fixture as the objects the oracle / product expect."""

    def __init__(self, g):
        self.num = g["ds/num"].astype(np.int64)
        self.num_list = np.cumsum(self.num)
        self.n_cells = int(self.num[0])
        self.bins = [int(b) for b in self.num[1:]]
        bins = np.array(self.bins, dtype=np.int32)
        bin_off = np.concatenate([[0], np.cumsum(bins)]).astype(np.int64)
        self.csr = PackedCSR(self.n_cells, bins, bin_off, g["ds/row_ptr"], g["ds/col"], g["ds/val"])
        self.attribute_dict = g["ds/attribute_dict"]
        self.cell_feats = g["ds/cell_feats"]
        self.cell_embed_feats = g["ds/cell_embed_feats"]
        self.bin_feats = [g["ds/bin_feats_%d" % j] for j in range(len(self.bins))]
        self.feats = [self.cell_embed_feats] + self.bin_feats

    def matrices(self, chrom, cell_1based):
        return self.csr.matrix(int(chrom), int(cell_1based) - 1)
