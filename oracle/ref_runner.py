"""Drives the UNMODIFIED reference (Impute.impute_process, Impute.py:123-302) on a synthetic dataset -- test infrastructure.

`bench.py --impl reference` and the `cpu_baseline` leg call this when the reference's own modules are importable (the build
container's /root/reference, or the `pip install --target baseline/_ref` copy that travels to the GPU box); otherwise they time the
oracle port (oracle/impute.py) and say so (`cpu_baseline.kind`).  Nothing in higashi_b200/ imports this file.

What runs is the reference's code, unchanged: `MultipleEmbedding`, `Hyper_SAGNN`, `GraphSageEncoder_with_weights` are built by the
reference's constructors (random init), the contacts / kNN lists / config are written in the formats `impute_process` reads
(`sparse_nondiag_adj_nbr_1.npy` object array, `weighted_info.npy`, `config.json`, `node_feats.hdf5`), and one call of
`impute_process(config, model, name, mode, cell_start, cell_end, sparse_path, weighted_info)` imputes a cell range: `prep_one`
-> `fix_cell2` -> `predict(batch 1e5)` -> one `cell_%d` dataset per chromosome.  The only substitutions are the import shims
(oracle/ref_shims.py): a dict-backed `h5py` (libhdf5 is not in the image; the gzip-6 compression of the real sink is therefore NOT
in the reference's time, which favours the reference), a stub `cooler`, and the `get_csr_submatrix` alias newer scipy dropped.
"""
import json
import os
import tempfile
import time

import numpy as np


def available():
    from . import ref_shims
    return os.path.isfile(os.path.join(ref_shims.REFERENCE_ROOT, "higashi", "Impute.py"))


class ReferenceImputer:
    """One synthetic dataset + one reference model, ready for repeated `impute(cell_start, cell_end)` calls."""

    def __init__(self, ds, nbr, w, mode="classification", seed=3):
        import torch
        from . import ref_shims
        self.shims = ref_shims
        RM, RI, _ = ref_shims.import_reference()
        self.RI, self.RM, self.ds, self.mode = RI, RM, ds, mode
        d = ds.d
        torch.manual_seed(seed)
        emb = [ds.cell_embed_feats] + list(ds.bin_feats)
        nei = RM.MultipleEmbedding(emb, d, False, ds.num_list, emb)                       # Higashi_wrapper.py:812-819
        model = RM.Hyper_SAGNN(n_head=8, d_model=d, d_k=16, d_v=16, diag_mask=True, bottle_neck=d,
                               attribute_dict=ds.attribute_dict, cell_feats=ds.cell_feats,
                               encoder_dynamic_nn=nei, encoder_static_nn=nei, chrom_num=len(ds.bins))     # :835-848
        nei.off_hook([0])                                                                 # :1384
        n_nodes = int(ds.num_list[-1])
        start_end = np.zeros((n_nodes + 1, 2), dtype="int")
        for j in range(len(ds.bins)):
            start_end[ds.num_list[j] + 1: ds.num_list[j + 1] + 1] = (ds.num_list[j], ds.num_list[j + 1])
        start_end[1: ds.n_cells + 1] = (0, ds.n_cells)
        enc = RM.GraphSageEncoder_with_weights(features=nei, linear_features=nei, feature_dim=d, embed_dim=d, num_sample=8,
                                               gcn=False, num_list=ds.num_list, transfer_range=0, start_end_dict=start_end,
                                               pass_pseudo_id=False, remove=True, pass_remove=False)      # :1415-1424
        model.encode1.dynamic_nn = enc
        self.model = model
        self.tmp = tempfile.mkdtemp(prefix="hsg_ref_")
        self.chrom_list = ["chr%d" % (j + 1) for j in range(len(ds.bins))]
        res = 1000000
        cfg = dict(resolution=res, impute_list=self.chrom_list, chrom_list=self.chrom_list, temp_dir=self.tmp,
                   minimum_distance=res * max(int(ds.min_bin), 0) if ds.min_bin >= 0 else -1,
                   maximum_distance=-1 if ds.max_bin < 0 else res * int(ds.max_bin), local_transfer_range=0,
                   embedding_name="emb", impute_verbose=0)
        self.cfg_path = os.path.join(self.tmp, "config.json")
        json.dump(cfg, open(self.cfg_path, "w"))
        ref_shims.H5_STORE[os.path.join(self.tmp, "node_feats.hdf5")] = {"num": ds.num}
        os.makedirs(os.path.join(self.tmp, "embed"))                                      # get_weights reads it (Impute.py:62-76)
        rng = np.random.Generator(np.random.PCG64(5))
        np.save(os.path.join(self.tmp, "embed", "emb_0_origin.npy"), rng.standard_normal((ds.n_cells, 8)).astype(np.float32))
        self.sparse_path = os.path.join(self.tmp, "sparse_nondiag_adj_nbr_1.npy")
        np.save(self.sparse_path, ds.csr.to_object_array(), allow_pickle=True)
        self.winfo = None
        self.k = 0
        if nbr is not None and nbr.shape[1] > 1:
            # weighted_info.npy = [cell_neighbor_list, weight_dict] (Higashi_wrapper.py:1303-1345): 1-based cell ids, self first
            self.k = nbr.shape[1] - 1
            nl = [[]] + [[int(c) for c in row] for row in nbr]
            wd = {}
            for i in range(1, len(nl)):
                for c, ww in zip(nl[i], w[i - 1]):
                    wd[(c, i)] = float(ww)
            self.winfo = os.path.join(self.tmp, "weighted_info.npy")
            np.save(self.winfo, np.array([nl, wd], dtype=object), allow_pickle=True)
        self.name = "emb_nbr_%d_impute" % self.k
        self.P = None

    def set_threads(self, n):
        import torch
        torch.set_num_threads(int(n))          # after the import: Impute.py:14 pins the process to one thread at import time

    def impute(self, cell_start, cell_end):
        """One reference call; returns (seconds, triplets written)."""
        import contextlib
        import io
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            self.RI.impute_process(self.cfg_path, self.model, self.name, self.mode, int(cell_start), int(cell_end),
                                   self.sparse_path, self.winfo)
        dt = time.perf_counter() - t0
        n = 0
        for chrom in self.chrom_list:
            f = self.shims.H5_STORE[os.path.join(self.tmp, "%s_%s.hdf5" % (chrom, self.name))]
            n += sum(len(f["cell_%d" % c]) for c in range(cell_start, cell_end))
        return dt, n

    def scores(self, cell):
        """The imputed vector of one cell of the LAST call, chromosomes concatenated (parity checks)."""
        return np.concatenate([np.asarray(self.shims.H5_STORE[os.path.join(self.tmp, "%s_%s.hdf5" % (c, self.name))]["cell_%d" % cell])
                               for c in self.chrom_list])

    def state_arrays(self):
        out = {}
        for k, v in self.model.state_dict().items():
            if k.startswith(("encode1.dynamic_nn.features.", "encode1.dynamic_nn.linear_features.",
                             "encode1.dynamic_nn.aggregator.", "encode1.dynamic_nn.linear_aggregator.")):
                continue
            out[k] = v.detach().cpu().numpy()
        return out
