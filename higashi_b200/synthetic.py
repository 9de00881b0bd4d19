"""Synthetic scHi-C datasets of the shapes BASELINE.json names.

Everything the hot path consumes is produced here directly in the layouts the
reference's pre-processing (Process.py, not on the hot path) would have left in
``temp_dir``:

* per-(chromosome, cell) contact matrices -- symmetric, zero diagonal, fp32 CSR,
  distance decay P(|i-j| = s) ~ 1/s, counts 1 + Poisson(0.3), empty rows get a
  unit diagonal, scaled by the per-cell read normaliser (reference
  Process.py:377-383);
* bin node features = L1-row-normalised pseudo-bulk, globally standardised,
  followed by a one-hot chromosome block (Process.py:580-590,
  Higashi_wrapper.py:634-646);
* cell node features N(0,1) of width min(0.5 n_cells, 2400) (Process.py:749-753);
* the attribute matrix and cell covariate (Higashi_wrapper.py:649-663, 763-780).

All randomness comes from ``numpy.random.Generator(PCG64(seed))``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

HG19 = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663,
        146364022, 141213431, 135534747, 135006516, 133851895, 115169878, 107349540,
        102531392, 90354753, 81195210, 78077248, 59128983, 63025520, 48129895, 51304566]
MM10 = [195471971, 182113224, 160039680, 156508116, 151834684, 149736546, 145441459,
        129401213, 124595110, 130694993, 122082543, 120129022, 120421639, 124902244,
        104043685, 98207768, 94987271, 90702639, 61431566]


def chrom_bins(genome: str, res: int, chroms: Optional[List[int]] = None) -> List[int]:
    """Number of bins per chromosome = ceil(length / resolution)."""
    lens = {"hg19": HG19, "mm10": MM10}[genome]
    if chroms is not None:
        lens = [lens[c] for c in chroms]
    return [int(-(-l // res)) for l in lens]


# The five BASELINE.json configurations.  `k` = neighbour cells.
CONFIGS = {
    "C1": dict(n_cells=620, genome="hg19", res=1_000_000, chroms=[0], density=0.017, k=0,
               d=64, min_bin=1, max_bin=-1),
    "C2": dict(n_cells=6000, genome="hg19", res=1_000_000, chroms=None, density=0.013, k=4,
               d=64, min_bin=1, max_bin=-1),
    "C3": dict(n_cells=4200, genome="hg19", res=500_000, chroms=None, density=0.05, k=4,
               d=64, min_bin=1, max_bin=-1),
    "C4": dict(n_cells=2000, genome="mm10", res=500_000, chroms=None, density=0.08, k=4,
               d=64, min_bin=1, max_bin=-1),
    "C5": dict(n_cells=4000, genome="hg19", res=50_000, chroms=[1], density=0.003, k=4,
               d=64, min_bin=1, max_bin=200),
}


@dataclass
class PackedCSR:
    """All (cell, chromosome) contact matrices as ONE CSR blob.

    Row id of (cell c, chromosome j, local bin b) is ``c * n_bins + bin_off[j] + b``;
    column indices are local to the chromosome.  This is also the device layout.
    """
    n_cells: int
    bins: np.ndarray          # int32 [C]
    bin_off: np.ndarray       # int64 [C+1]
    row_ptr: np.ndarray       # int64 [n_cells * n_bins + 1]
    col: np.ndarray           # int32 [nnz]
    val: np.ndarray           # float32 [nnz]

    @property
    def n_bins(self) -> int:
        return int(self.bin_off[-1])

    def matrix(self, chrom: int, cell: int):
        """scipy CSR of one (chromosome, cell) -- the reference's object format."""
        from scipy.sparse import csr_matrix
        n = int(self.bins[chrom])
        r0 = cell * self.n_bins + int(self.bin_off[chrom])
        ptr = self.row_ptr[r0:r0 + n + 1]
        lo, hi = int(ptr[0]), int(ptr[-1])
        return csr_matrix((self.val[lo:hi].copy(), self.col[lo:hi].copy(),
                           (ptr - lo).astype(np.int32)), shape=(n, n))

    def to_object_array(self):
        """``sparse_nondiag_adj_nbr_1.npy`` layout: object array [C][n_cells] of CSR."""
        out = np.empty((len(self.bins), self.n_cells), dtype=object)
        for j in range(len(self.bins)):
            for c in range(self.n_cells):
                out[j, c] = self.matrix(j, c)
        return out

    @staticmethod
    def from_object_array(sparse_chrom_list) -> "PackedCSR":
        """Pack the reference's ``[C][n_cells]`` scipy-CSR object array."""
        C = len(sparse_chrom_list)
        n_cells = len(sparse_chrom_list[0])
        bins = np.array([sparse_chrom_list[j][0].shape[0] for j in range(C)], dtype=np.int32)
        bin_off = np.concatenate([[0], np.cumsum(bins)]).astype(np.int64)
        n_bins = int(bin_off[-1])
        counts = np.zeros(n_cells * n_bins, dtype=np.int64)
        cols, vals = [], []
        for c in range(n_cells):
            for j in range(C):
                m = sparse_chrom_list[j][c].tocsr()
                m.sort_indices()
                r0 = c * n_bins + int(bin_off[j])
                counts[r0:r0 + bins[j]] = np.diff(m.indptr)
                cols.append(m.indices.astype(np.int32))
                vals.append(m.data.astype(np.float32))
        row_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        col = np.concatenate(cols) if cols else np.zeros(0, np.int32)
        val = np.concatenate(vals) if vals else np.zeros(0, np.float32)
        return PackedCSR(n_cells, bins, bin_off, row_ptr, col, val)

    # ---- stable binary layout: replaces unpickling sparse_nondiag_adj_nbr_1.npy on every launch -------------
    # header: magic "HSGCSR01", int64 n_cells, C, n_rows, nnz ; then bins int32[C] (padded to 8 bytes), row_ptr int64[n_rows+1],
    # col int32[nnz] (padded to 8 bytes), val float32[nnz].  All little-endian, every array 8-byte aligned, so the file can be
    # np.memmap'ed and handed to hsg_set_contacts without a copy on the host.
    MAGIC = b"HSGCSR01"

    def save_blob(self, path: str):
        n_rows, nnz, C = len(self.row_ptr) - 1, len(self.col), len(self.bins)
        pad = lambda n_bytes: (-n_bytes) % 8
        with open(path, "wb") as f:
            f.write(self.MAGIC)
            f.write(np.array([self.n_cells, C, n_rows, nnz], dtype="<i8").tobytes())
            f.write(np.ascontiguousarray(self.bins, dtype="<i4").tobytes()); f.write(b"\0" * pad(4 * C))
            f.write(np.ascontiguousarray(self.row_ptr, dtype="<i8").tobytes())
            f.write(np.ascontiguousarray(self.col, dtype="<i4").tobytes()); f.write(b"\0" * pad(4 * nnz))
            f.write(np.ascontiguousarray(self.val, dtype="<f4").tobytes())

    @staticmethod
    def from_npy_cached(npy_path: str) -> "PackedCSR":
        """Loads `<npy_path>.hsgcsr` if it is at least as new as the pickled object array, else packs the array once and writes
        the blob next to it (falls back silently to no caching on a read-only directory)."""
        import os
        blob = npy_path + ".hsgcsr"
        try:
            if os.path.exists(blob) and os.path.getmtime(blob) >= os.path.getmtime(npy_path):
                return PackedCSR.load_blob(blob)
        except (OSError, ValueError):
            pass
        csr = PackedCSR.from_object_array(np.load(npy_path, allow_pickle=True))
        # several workers may reach this point together (impute_multi_gpu starts one per GPU) while others already memmap the
        # blob: write a private temporary next to it and rename it into place (atomic on POSIX), never truncate a live file
        tmp = "%s.tmp.%d" % (blob, os.getpid())
        try:
            csr.save_blob(tmp)
            os.replace(tmp, blob)
        except OSError:
            try:
                os.remove(tmp)
            except OSError:
                pass
        return csr

    @staticmethod
    def load_blob(path: str, mmap: bool = True) -> "PackedCSR":
        with open(path, "rb") as f:
            head = f.read(8 + 32)
        if head[:8] != PackedCSR.MAGIC:
            raise ValueError("%s is not a packed contact blob (bad magic)" % path)
        n_cells, C, n_rows, nnz = (int(v) for v in np.frombuffer(head[8:], dtype="<i8"))
        pad = lambda n_bytes: (-n_bytes) % 8
        off = 40
        take = (lambda dt, n, o: np.memmap(path, dtype=dt, mode="r", offset=o, shape=(n,))) if mmap else \
            (lambda dt, n, o: np.fromfile(path, dtype=dt, count=n, offset=o))
        bins = np.array(take("<i4", C, off)); off += 4 * C + pad(4 * C)
        row_ptr = take("<i8", n_rows + 1, off); off += 8 * (n_rows + 1)
        col = take("<i4", nnz, off) if nnz else np.zeros(0, np.int32); off += 4 * nnz + pad(4 * nnz)
        val = take("<f4", nnz, off) if nnz else np.zeros(0, np.float32)
        if int(bins.sum()) * n_cells != n_rows or int(row_ptr[0]) != 0 or int(row_ptr[-1]) != nnz:
            raise ValueError("%s: inconsistent header / row_ptr" % path)
        bin_off = np.concatenate([[0], np.cumsum(bins)]).astype(np.int64)
        return PackedCSR(n_cells, bins.astype(np.int32), bin_off, row_ptr, col, val)


@dataclass
class SynthDataset:
    name: str
    n_cells: int
    bins: List[int]
    d: int
    k: int
    min_bin: int
    max_bin: int
    csr: PackedCSR
    bin_feats: List[np.ndarray] = field(default_factory=list)   # [n_c, w_c + C] each
    cell_embed_feats: Optional[np.ndarray] = None                # [n_cells, F]
    attribute_dict: Optional[np.ndarray] = None                  # [n_cells+1+n_bins, C+1]
    cell_feats: Optional[np.ndarray] = None                      # [n_cells+1, 1]

    @property
    def num(self) -> np.ndarray:
        return np.array([self.n_cells] + list(self.bins), dtype=np.int64)

    @property
    def num_list(self) -> np.ndarray:
        return np.cumsum(self.num)


def _decay_cdf(n: int) -> np.ndarray:
    s = np.arange(1, n, dtype=np.float64)
    p = 1.0 / s
    return np.cumsum(p / p.sum())


def generate_contacts(n_cells: int, bins: List[int], density: float, seed: int,
                      dtype=np.float32) -> PackedCSR:
    """Vectorised generator of the packed per-cell contact matrices."""
    rng = np.random.Generator(np.random.PCG64(seed))
    C = len(bins)
    bins_a = np.asarray(bins, dtype=np.int32)
    bin_off = np.concatenate([[0], np.cumsum(bins_a)]).astype(np.int64)
    n_bins = int(bin_off[-1])
    rows_all, cols_all, vals_all = [], [], []
    for j, n in enumerate(bins):
        # expected nnz (both triangles) = density * n^2  ->  contacts in the upper triangle
        m_target = max(1.0, density * n * n / 2.0)
        # draw ~15 % more to compensate for rejected (out-of-range) and merged draws
        m_draw = rng.poisson(m_target * 1.25, size=n_cells).astype(np.int64)
        M = int(m_draw.sum())
        cell = np.repeat(np.arange(n_cells, dtype=np.int64), m_draw)
        i = rng.integers(0, n - 1, size=M, dtype=np.int64)
        cdf = _decay_cdf(n)
        s = np.searchsorted(cdf, rng.random(M), side="right").astype(np.int64) + 1
        jj = i + s
        cnt = 1.0 + rng.poisson(0.3, size=M)
        ok = jj < n
        cell, i, jj, cnt = cell[ok], i[ok], jj[ok], cnt[ok]
        key = (cell * n + i) * n + jj
        ukey, inv = np.unique(key, return_inverse=True)
        w = np.bincount(inv, weights=cnt, minlength=len(ukey))
        ucell = ukey // (n * n)
        ui = (ukey // n) % n
        uj = ukey % n
        # per-cell read normaliser: mean reads per cell / this cell's reads (Process.py:379,549)
        reads = np.bincount(ucell, weights=w, minlength=n_cells)
        per_cell_read = reads.sum() / n_cells
        scale = per_cell_read / (reads + 1e-15)
        # symmetrise
        r = np.concatenate([ui, uj])
        c = np.concatenate([uj, ui])
        cc = np.concatenate([ucell, ucell])
        v = np.concatenate([w, w])
        # unit diagonal on empty rows (Process.py:382), before the scaling
        has = np.zeros(n_cells * n, dtype=bool)
        has[cc * n + r] = True
        empty = np.flatnonzero(~has)
        if len(empty):
            ecell, erow = empty // n, empty % n
            r = np.concatenate([r, erow]); c = np.concatenate([c, erow])
            cc = np.concatenate([cc, ecell]); v = np.concatenate([v, np.ones(len(empty))])
        v = v * scale[cc]
        rows_all.append(cc * n_bins + bin_off[j] + r)
        cols_all.append(c.astype(np.int32))
        vals_all.append(v.astype(dtype))
    rows = np.concatenate(rows_all)
    cols = np.concatenate(cols_all)
    vals = np.concatenate(vals_all)
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    counts = np.bincount(rows, minlength=n_cells * n_bins)
    row_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return PackedCSR(n_cells, bins_a, bin_off, row_ptr, cols, vals.astype(np.float32))


def pseudo_bulk_features(csr: PackedCSR) -> List[np.ndarray]:
    """Bin node features per chromosome (see module docstring)."""
    C = len(csr.bins)
    n_bins = csr.n_bins
    nnz_row = np.diff(csr.row_ptr)
    grow = np.repeat(np.arange(csr.n_cells * n_bins, dtype=np.int64), nnz_row)
    lrow = grow % n_bins                       # row inside the genome
    feats = []
    for j in range(C):
        n = int(csr.bins[j])
        lo, hi = int(csr.bin_off[j]), int(csr.bin_off[j + 1])
        sel = (lrow >= lo) & (lrow < hi)
        flat = (lrow[sel] - lo) * n + csr.col[sel]
        bulk = np.bincount(flat, weights=csr.val[sel].astype(np.float64), minlength=n * n)
        bulk = bulk.reshape(n, n)
        bulk = bulk + np.diag((bulk.sum(-1) == 0).astype(np.float64))
        mean_, std_ = bulk.mean(), bulk.std()
        np.clip(bulk, None, mean_ + 10 * std_, out=bulk)
        bulk = bulk / np.maximum(np.abs(bulk).sum(-1, keepdims=True), 1e-30)
        if n >= 3000:                           # Process.py:615-622 column down-sampling
            sf = int(round(n / 3000))
            w = (n // sf) * sf
            bulk = bulk[:, :w].reshape(n, w // sf, sf).mean(-1)
        bulk = (bulk - bulk.mean()) / (bulk.std() + 1e-30)   # wrapper:638 global standardise
        onehot = np.zeros((n, C), dtype=np.float32)
        onehot[:, j] = 1
        feats.append(np.concatenate([bulk.astype(np.float32), onehot], axis=-1))
    return feats


def make_attributes(n_cells: int, bins: List[int]):
    """attribute_dict (wrapper:649-663) and cell_feats (wrapper:763-780)."""
    C = len(bins)
    blocks = []
    for j, n in enumerate(bins):
        chrom = np.zeros((n, C)); chrom[:, j] = 1
        coor = (np.arange(n).reshape(-1, 1).astype(np.float32)) / bins[0]
        blocks.append(np.concatenate([chrom, coor], axis=-1))
    attr = np.concatenate(blocks, axis=0)
    attr = np.concatenate([np.zeros((n_cells + 1, C + 1)), attr], axis=0).astype(np.float32)
    return attr


def make_dataset(name: str = "C1", *, n_cells: Optional[int] = None, bins: Optional[List[int]] = None,
                 seed: Optional[int] = None, shard: int = 0, with_features: bool = True,
                 **overrides) -> SynthDataset:
    """Build one of the named configurations (optionally shrunk for tests)."""
    cfg = dict(CONFIGS[name]) if name in CONFIGS else dict(CONFIGS["C1"])
    cfg.update(overrides)
    if n_cells is not None:
        cfg["n_cells"] = n_cells
    if bins is None:
        bins = chrom_bins(cfg["genome"], cfg["res"], cfg["chroms"])
    idx = int(name[1:]) if name in CONFIGS else 0
    if seed is None:
        seed = 1000 * idx + shard                     # SURVEY 8d seeding rule
    csr = generate_contacts(cfg["n_cells"], bins, cfg["density"], seed)
    ds = SynthDataset(name=name, n_cells=cfg["n_cells"], bins=list(bins), d=cfg["d"], k=cfg["k"],
                      min_bin=cfg["min_bin"], max_bin=cfg["max_bin"], csr=csr)
    if with_features:
        rng = np.random.Generator(np.random.PCG64(seed + 7919))
        ds.bin_feats = pseudo_bulk_features(csr)
        F = int(min(max(4, 0.5 * ds.n_cells), 2400))
        ds.cell_embed_feats = rng.standard_normal((ds.n_cells, F)).astype(np.float32)
        ds.attribute_dict = make_attributes(ds.n_cells, ds.bins)
        reads = np.add.reduceat(csr.val.astype(np.float64),
                                np.minimum(csr.row_ptr[:-1:csr.n_bins], len(csr.val) - 1))
        cf = np.log1p(reads.reshape(-1, 1))
        ds.cell_feats = np.concatenate([np.zeros((1, 1)), cf], axis=0).astype(np.float32)
    return ds


def random_state_dict(ds: "SynthDataset", d: int, seed: int = 0):
    """Random-init weights with the reference's state_dict names and shapes (Modules.py constructors:
    Linear/Conv1d default init U(-1/sqrt(fan_in), 1/sqrt(fan_in)); attention projections N(0, 2/(d+16)),
    Modules.py:998-1003; LayerNorm affine perturbed so that no term is an identity).
    Returns (sd, feats) with feats = [cell features, bin features per chromosome]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    C = len(ds.bins)
    xin = 2 * (C + 1) + ds.cell_feats.shape[-1]

    def U(*shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return rng.uniform(-b, b, size=shape).astype(np.float32)

    sd = {}
    for nm in ("w_qs", "w_ks", "w_vs"):
        sd["encode1.mul_head_attn.%s.weight" % nm] = (rng.standard_normal((128, d)) * np.sqrt(2.0 / (d + 16))).astype(np.float32)
    sd["encode1.mul_head_attn.fc1.w_stack.0.weight"] = U(d, 128, fan_in=128)
    sd["encode1.mul_head_attn.fc2.w_stack.0.weight"] = U(d, 128, fan_in=128)
    for nm in ("encode1.mul_head_attn.layer_norm1", "encode1.mul_head_attn.layer_norm2", "encode1.mul_head_attn.layer_norm3",
               "encode1.pff_n1.layer_norm", "layer_norm1", "layer_norm2"):
        sd[nm + ".weight"] = (1 + 0.2 * rng.standard_normal(d)).astype(np.float32)
        sd[nm + ".bias"] = (0.1 * rng.standard_normal(d)).astype(np.float32)
    for i in (0, 1):
        sd["encode1.pff_n1.w_stack.%d.weight" % i] = U(d, d, 1, fan_in=d)
        sd["encode1.pff_n1.w_stack.%d.bias" % i] = U(d, fan_in=d)
    for head in ("pff_classifier", "pff_classifier_var", "pff_classifier_proba"):
        sd[head + ".w_stack.0.weight"] = U(d // 2, d, 1, fan_in=d)
        sd[head + ".w_stack.0.bias"] = U(d // 2, fan_in=d)
        sd[head + ".w_stack.1.weight"] = U(1, d // 2, 1, fan_in=d // 2)
        sd[head + ".w_stack.1.bias"] = U(1, fan_in=d // 2)
    sd["encode1.dynamic_nn.nn.weight"] = U(d, 2 * d, fan_in=2 * d)
    sd["encode1.dynamic_nn.nn.bias"] = U(d, fan_in=2 * d)
    for xp in ("extra_proba", "extra_proba2", "extra_proba3"):
        sd[xp + ".w_stack.0.weight"] = U(4, xin, fan_in=xin)
        sd[xp + ".w_stack.0.bias"] = U(4, fan_in=xin)
        sd[xp + ".w_stack.1.weight"] = U(1, 4, fan_in=4)
        sd[xp + ".w_stack.1.bias"] = U(1, fan_in=4)
    sd["attribute_dict_embedding.weight"] = ds.attribute_dict
    feats = [ds.cell_embed_feats] + list(ds.bin_feats)
    for i, f in enumerate(feats):
        sd["encode1.static_nn.wstack.%d.weight_list.0" % i] = U(d, f.shape[1], fan_in=f.shape[1])
        sd["encode1.static_nn.wstack.%d.bias_list.0" % i] = U(d, fan_in=f.shape[1])
    return sd, feats


def write_temp_dir(ds: "SynthDataset", temp_dir: str, chrom_names=None, test_frac: float = 0.1, seed: int = 0):
    """Write the files the reference's Process.py leaves in `temp_dir` (its output = the hot path's input contract):
    node_feats (num, train/test hyperedges + weights per chromosome, sparsity, start_end_dict, extra_cell_feats, cell2weight,
    bin features "{c}", cell features "cell/{c}"; Process.py:627,773,781-858) as `.npz` (h5py is absent here),
    `chrom_start_end.npy` and `sparse_nondiag_adj_nbr_1.npy` (object array [C][n_cells] of scipy CSR, Process.py:783-784)."""
    import os
    rng = np.random.Generator(np.random.PCG64(seed))
    os.makedirs(temp_dir, exist_ok=True)
    C = len(ds.bins)
    chrom_names = chrom_names or ["chr%d" % (j + 1) for j in range(C)]
    out = {"num": ds.num}
    csr, nl = ds.csr, ds.num_list
    start_end = np.zeros((int(nl[-1]), 2), dtype=np.int64)
    start_end[:ds.n_cells] = (0, ds.n_cells)
    reads = np.zeros((ds.n_cells, C))
    off, cse = 0, []
    for j, name in enumerate(chrom_names):
        n = ds.bins[j]
        start_end[nl[j]:nl[j + 1]] = (nl[j], nl[j + 1])
        cse.append((off, off + n)); off += n
        trip, wts = [], []
        for c in range(ds.n_cells):
            m = csr.matrix(j, c).tocoo()
            up = (m.col - m.row) >= 1
            reads[c, j] = m.data[up].sum()
            if up.any():
                trip.append(np.stack([np.full(up.sum(), c + 1), nl[j] + m.row[up] + 1, nl[j] + m.col[up] + 1], axis=-1))
                wts.append(np.maximum(1.0, np.round(m.data[up])))
        trip = np.concatenate(trip).astype(np.int64); wts = np.concatenate(wts).astype(np.float32)
        perm = rng.permutation(len(trip))
        n_test = max(1, int(test_frac * len(trip)))
        for tag, idx in (("test", perm[:n_test]), ("train", perm[n_test:])):
            out["%s_data_%s" % (tag, name)] = trip[idx]
            out["%s_weight_%s" % (tag, name)] = wts[idx]
            out["%s_chrom_%s" % (tag, name)] = np.full(len(idx), j, dtype=np.int8)
        out["%d" % j] = ds.bin_feats[j][:, :-C] if ds.bin_feats else np.eye(n, dtype=np.float32)
        F = ds.cell_embed_feats.shape[1]
        w = max(1, F // C)
        out["cell/%d" % j] = ds.cell_embed_feats[:, j * w:(j + 1) * w if j < C - 1 else F]
    out["sparsity"] = np.array([(csr.row_ptr[(c + 1) * csr.n_bins] - csr.row_ptr[c * csr.n_bins]) / float(sum(b * b for b in ds.bins))
                                for c in range(ds.n_cells)])
    out["start_end_dict"] = start_end
    out["extra_cell_feats"] = reads
    out["cell2weight"] = (reads.sum(-1) / max(1e-9, reads.sum(-1).mean())).astype(np.float32)
    out["distance2weight"] = np.ones(10, dtype=np.float32)
    np.savez(os.path.join(temp_dir, "node_feats.npz"), **out)
    np.save(os.path.join(temp_dir, "chrom_start_end.npy"), np.array(cse, dtype=np.int64))
    np.save(os.path.join(temp_dir, "sparse_nondiag_adj_nbr_1.npy"), csr.to_object_array(), allow_pickle=True)
    return chrom_names


def make_dataset_cells(name: str, cells, *, block: int = 64, threads: int = 4, seed: Optional[int] = None) -> "SynthDataset":
    """ONE dataset of the named configuration, materialised for a SUBSET of its cells (strong scaling: a rank holds its own cell
    range plus the kNN halo).  The dataset is defined block-wise -- cells [b * block, (b + 1) * block) come from
    generate_contacts(seed = base + 7919 * (b + 1)) -- so every rank that needs a cell generates the same contacts for it without
    generating the other blocks.  Returns a SynthDataset over the selected cells in ascending order of their global ids (local id =
    position in `cells`), `ds.global_cells` holds the global ids, `ds.n_cells_global` the size of the full dataset."""
    from concurrent.futures import ThreadPoolExecutor
    cfg = dict(CONFIGS[name])
    bins = chrom_bins(cfg["genome"], cfg["res"], cfg["chroms"])
    base = 1000 * int(name[1:]) if seed is None else seed
    cells = np.unique(np.asarray(cells, dtype=np.int64))
    n_glob = cfg["n_cells"]
    assert len(cells) and cells[0] >= 0 and cells[-1] < n_glob
    blocks = np.unique(np.concatenate([[0], cells // block]))     # block 0 always: the bin features are defined on it

    def one(b):
        nb = int(min(block, n_glob - b * block))
        return int(b), generate_contacts(nb, bins, cfg["density"], base + 7919 * (int(b) + 1))
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        parts = dict(ex.map(one, blocks))
    n_bins = int(sum(bins))
    rp_parts, col_parts, val_parts, off = [np.zeros(1, dtype=np.int64)], [], [], 0
    for c in cells:
        blk = parts[int(c // block)]
        q = int(c % block)
        lo, hi = int(blk.row_ptr[q * n_bins]), int(blk.row_ptr[(q + 1) * n_bins])
        rp_parts.append(blk.row_ptr[q * n_bins + 1:(q + 1) * n_bins + 1] - lo + off)
        col_parts.append(blk.col[lo:hi]); val_parts.append(blk.val[lo:hi])
        off += hi - lo
    any_blk = next(iter(parts.values()))
    csr = PackedCSR(len(cells), any_blk.bins, any_blk.bin_off, np.concatenate(rp_parts), np.concatenate(col_parts), np.concatenate(val_parts))
    ds = SynthDataset(name=name, n_cells=len(cells), bins=list(bins), d=cfg["d"], k=cfg["k"], min_bin=cfg["min_bin"],
                      max_bin=cfg["max_bin"], csr=csr)
    rng = np.random.Generator(np.random.PCG64(base + 7919))
    ds.bin_feats = pseudo_bulk_features(parts[0])                 # the same on every rank, whatever cells it holds
    F = int(min(max(4, 0.5 * n_glob), 2400))
    ds.cell_embed_feats = rng.standard_normal((n_glob, F)).astype(np.float32)[cells]
    ds.attribute_dict = make_attributes(ds.n_cells, ds.bins)
    reads = np.add.reduceat(csr.val.astype(np.float64), np.minimum(csr.row_ptr[:-1:csr.n_bins], len(csr.val) - 1))
    ds.cell_feats = np.concatenate([np.zeros((1, 1)), np.log1p(reads.reshape(-1, 1))], axis=0).astype(np.float32)
    ds.global_cells = cells
    ds.n_cells_global = n_glob
    return ds
