"""Mirror of the reference orchestrator `higashi/Higashi_wrapper.py` (class `Higashi`) for the scorer hot path.

Same public methods and config-JSON keys as the reference (Higashi_wrapper.py:439-1641):
    Higashi(config_path).prep_model() -> train_for_embeddings(max_epochs) -> train_for_imputation_nbr_0() / _no_nbr()
    -> impute_no_nbr() -> train_for_imputation_with_nbr() -> impute_with_nbr();  fetch_cell_embeddings(), fetch_map(chrom, cell)
What runs where: batch bookkeeping (DataGenerator, negative sampling, neighbour lists) is host numpy -- the reference does it
in CPU worker processes; every tensor op of the scorer (forward, loss, backward, imputation, kNN graph) runs in libhsg on the
GPU.  `process_data()` (the Process.py ETL) is out of scope: its output files are this class's input contract.

Files read from `temp_dir` (written by the reference's Process.py; `.npz` twins with the same keys are accepted because h5py is
not available in this image): node_feats.hdf5|npz, sparse_nondiag_adj_nbr_1.npy, chrom_start_end.npy.
Unsupported config options raise NotImplementedError: coassay, pre_cell_embed, batch_id / library_id / regress_cov,
correct_be_impute (they change the ETL-side feature construction, not the hot path)."""
from __future__ import annotations

import math
import os
import time

import numpy as np
import torch

from . import batches, parallel, sink
from . import engine as E
from .Higashi_backend import Modules as M
from .Higashi_backend.utils import get_config
from .Impute import impute_process
from .synthetic import PackedCSR
from .train_engine import LOSS_MODES, TrainEngine


class NegativeSampler:
    """Vectorised host restatement of generate_negative_cpu / check_nonzero (Higashi_wrapper.py:71-168): for every positive
    `neg_num` negatives by rejection sampling; hard negatives (probability pair_ratio) change one node, easy ones resample all
    three; a candidate is rejected while it touches a non-zero 3x3 neighbourhood of the cell's contact map, or when its bin
    distance is >= max_bin or <= 1; at most 50 trials.  Statistical, not bit, parity (the reference is unseeded)."""

    def __init__(self, csr: PackedCSR, num_list, max_bin: int, rng=None):
        self.csr, self.num_list, self.max_bin = csr, np.asarray(num_list), int(max_bin)
        self.rng = rng or np.random.default_rng()
        self._keys = {}

    def _chrom_keys(self, c):
        if c not in self._keys:
            csr, n = self.csr, int(self.csr.bins[c])
            cells = np.arange(csr.n_cells, dtype=np.int64)
            rows0 = cells[:, None] * csr.n_bins + int(csr.bin_off[c]) + np.arange(n)[None, :]
            lo, hi = csr.row_ptr[rows0.reshape(-1)], csr.row_ptr[rows0.reshape(-1) + 1]
            cnt = hi - lo
            rid = np.repeat(np.arange(len(cnt), dtype=np.int64), cnt)
            idx = np.arange(int(cnt.sum()), dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt) + np.repeat(lo, cnt)
            self._keys[c] = np.sort(rid * n + csr.col[idx])          # (cell*n + row)*n + col
        return self._keys[c]

    def _nonzero_near(self, c, cell0, b1, b2):
        n = int(self.csr.bins[c])
        keys = self._chrom_keys(c)
        hit = np.zeros(len(cell0), dtype=bool)
        for dr in (-1, 0, 1):
            for dc in (-1, 0, 1):
                r = np.clip(b1 + dr, 0, n - 2); cc = np.clip(b2 + dc, 0, n - 2)
                k = (cell0 * n + r) * n + cc
                pos = np.searchsorted(keys, k)
                hit |= (pos < len(keys)) & (keys[np.minimum(pos, len(keys) - 1)] == k)
        return hit

    def sample(self, x, chrom, neg_num, pair_ratio):
        rng, n, base = self.rng, int(self.csr.bins[chrom]), int(self.num_list[chrom]) + 1
        pos = np.repeat(np.asarray(x, dtype=np.int64), neg_num, axis=0)
        cand = pos.copy()
        done = np.zeros(len(cand), dtype=bool)
        hard = rng.random(len(cand)) <= pair_ratio
        change = rng.integers(0, 3, size=len(cand))
        for _ in range(50):
            bad = ~done & self._nonzero_near(chrom, cand[:, 0] - 1, cand[:, 1] - base, cand[:, 2] - base)
            done |= ~bad
            if not bad.any():
                break
            idx = np.flatnonzero(bad)
            m = len(idx)
            new = pos[idx].copy()
            h = hard[idx]
            ch = change[idx]
            # hard: change one node
            cell_new = rng.integers(0, self.csr.n_cells, size=m) + 1
            other = new[:, 2] - base
            lo = np.maximum(0, other - self.max_bin); hi = np.minimum(n, other + self.max_bin)
            bin_new = (lo + (rng.random(m) * np.maximum(hi - lo, 1)).astype(np.int64)) + base
            new[h & (ch == 0), 0] = cell_new[h & (ch == 0)]
            new[h & (ch == 1), 1] = bin_new[h & (ch == 1)]
            new[h & (ch == 2), 2] = bin_new[h & (ch == 2)]
            # easy: resample everything
            e = ~h
            b1 = rng.integers(0, n, size=m)
            lo = np.maximum(0, b1 - self.max_bin); hi = np.minimum(n, b1 + self.max_bin)
            b2 = lo + (rng.random(m) * np.maximum(hi - lo, 1)).astype(np.int64)
            new[e, 0] = cell_new[e]; new[e, 1] = b1[e] + base; new[e, 2] = b2[e] + base
            new[:, 1:] = np.sort(new[:, 1:], axis=1)
            dist = new[:, 2] - new[:, 1]
            ok = (dist < self.max_bin) & (dist > 1)
            new[~ok] = pos[idx][~ok]
            cand[idx] = new
        return cand[done]


class Higashi:
    def __init__(self, config_path):
        self.config_path = config_path
        self.config = get_config(config_path)
        for sub in ("", "embed", "model"):
            os.makedirs(os.path.join(self.config["temp_dir"], sub), exist_ok=True)
        self.update_num_per_training_epoch = 1000
        self.update_num_per_eval_epoch = 10
        self.cell_embeddings = None

    # ---- ETL is out of scope ---------------------------------------------------------------------------------------------
    def process_data(self, disable_mpl=False):
        raise NotImplementedError("process_data() is the reference's Process.py ETL (out of scope of the hot path); run it with the "
                                  "reference and point temp_dir at its outputs")

    # ---- config ----------------------------------------------------------------------------------------------------------
    def fetch_info_from_config(self):
        c = self.config
        for bad in ("coassay", "pre_cell_embed", "batch_id", "library_id", "correct_be_impute"):
            if c.get(bad):
                raise NotImplementedError("config option %r is not supported by the B200 hot-path mirror" % bad)
        if c.get("regress_cov"):
            raise NotImplementedError("regress_cov is not supported")
        if not torch.cuda.is_available():
            raise RuntimeError("the Higashi mirror needs a GPU (no CPU fallback)")
        self.gpu_num = c["gpu_num"]
        self.temp_dir, self.data_dir = c["temp_dir"], c.get("data_dir", c["temp_dir"])
        self.embed_dir = os.path.join(self.temp_dir, "embed")
        self.chrom_list, self.impute_list = c["chrom_list"], c["impute_list"]
        self.dimensions, self.res = c["dimensions"], c["resolution"]
        self.neighbor_num = c["neighbor_num"] + 1
        self.mode = c["loss_mode"]
        self.rank_thres = c.get("rank_thres", 1)
        self.embedding_name = c["embedding_name"]
        self.save_path = os.path.join(self.temp_dir, "model", "model.chkpt")
        self.embedding_epoch = c.get("embedding_epoch", 60)
        self.no_nbr_epoch = c.get("no_nbr_epoch", 45)
        self.with_nbr_epoch = c.get("with_nbr_epoch", 30)
        self.chrom_start_end = np.load(os.path.join(self.temp_dir, "chrom_start_end.npy"))

    # ---- model -----------------------------------------------------------------------------------------------------------
    def generate_attributes(self, nf):
        """Higashi_wrapper.py:594-665 (no batch-effect removal): globally standardised node features + one-hot chromosome,
        attribute matrix [n_nodes+1, C+1]."""
        std = lambda a: ((a - a.mean()) / (a.std() + 1e-30)).astype(np.float32)
        C = len(self.chrom_list)
        cell = np.concatenate([np.asarray(nf["cell/%d" % i]) for i in range(C)], axis=-1)
        if self.num[0] >= 10:
            emb_cell = std(cell)
            tgt_cell = std(np.concatenate([np.asarray(nf["cell/%d" % i]) for i in range(C)], axis=-1))
        else:
            emb_cell = tgt_cell = cell.astype(np.float32)
        embeddings, targets = [emb_cell], [tgt_cell]
        for i in range(C):
            t = std(np.asarray(nf["%d" % i]).astype(np.float32))
            chrom = np.zeros((len(t), C), dtype=np.float32); chrom[:, i] = 1
            embeddings.append(np.concatenate([t, chrom], axis=-1)); targets.append(embeddings[-1])
        attr = []
        for i in range(C):
            chrom = np.zeros((self.num[i + 1], C)); chrom[:, i] = 1
            coor = np.arange(self.num[i + 1]).reshape(-1, 1).astype(np.float32) / self.num[1]
            attr.append(np.concatenate([chrom, coor], axis=-1))
        attr = np.concatenate([np.zeros((self.num[0] + 1, C + 1)), np.concatenate(attr, axis=0)], axis=0).astype(np.float32)
        return embeddings, attr, targets

    def prep_model(self):
        self.fetch_info_from_config()
        nf = sink.read_node_feats(self.temp_dir)
        self.num = np.asarray(nf["num"]).astype(np.int64)
        self.num_list = np.cumsum(self.num)
        C = len(self.chrom_list)
        rd = lambda k: [np.asarray(nf["%s_%s" % (k, c)]) for c in self.chrom_list]
        train_data, test_data = [a.astype("int") for a in rd("train_data")], [a.astype("int") for a in rd("test_data")]
        train_weight, test_weight = [a.astype("float32") for a in rd("train_weight")], [a.astype("float32") for a in rd("test_weight")]
        train_chrom, test_chrom = [a.astype("int8") for a in rd("train_chrom")], [a.astype("int8") for a in rd("test_chrom")]
        self.total_sparsity_cell = np.asarray(nf["sparsity"])
        self.median_total_sparsity_cell = float(np.median(self.total_sparsity_cell))
        self.contractive_flag = self.median_total_sparsity_cell >= 0.05
        self.contractive_loss_weight = 1e-3 if self.contractive_flag else 0.0
        self.start_end_dict = np.concatenate([np.zeros((1, 2)), np.asarray(nf["start_end_dict"])], axis=0).astype("int")
        self.cell_feats1 = np.asarray(nf["cell2weight"]).astype(np.float32)
        cell_feats = np.log1p(np.sum(np.asarray(nf["extra_cell_feats"]), axis=-1, keepdims=True))
        self.cell_feats = np.concatenate([np.zeros((1, 1)), cell_feats], axis=0).astype(np.float32)
        # batch size and number of negatives (Higashi_wrapper.py:728, 748-759)
        self.batch_size = min(int(256 * max(1000000 / self.res, 1) * max(self.num[0] / 6000, 1)), 1280)
        total_possible = sum((e - s) * (e - s + 1) // 2 for s, e in self.chrom_start_end) * int(self.num[0])
        sparsity = sum(len(a) + len(b) for a, b in zip(train_data, test_data)) / max(1, total_possible)
        self.neg_num = 1 if sparsity > 0.3 else 2 if sparsity > 0.2 else 3 if sparsity > 0.15 else 4 if sparsity > 0.1 else 5
        self.batch_size *= (1 + self.neg_num)
        self.max_bin = int(np.max(self.num[1:]))
        if self.mode == "classification":
            mean_w = np.mean(np.concatenate(train_weight))
            tw = lambda ws: [np.minimum(np.log2(w + 1), np.quantile(np.log2(w + 1), 0.99)) / mean_w * self.neg_num if len(w) else w for w in ws]
            train_weight, test_weight = tw(train_weight), tw(test_weight)       # transform_weight_class, utils.py:40-44
        elif self.mode == "rank":
            train_weight, test_weight = [w + 1 for w in train_weight], [w + 1 for w in test_weight]
        embeddings, attr, targets = self.generate_attributes(nf)
        self.feats = embeddings
        # packed binary blob next to the pickled object array: unpickling thousands of scipy matrices happens once, not per launch
        self.csr = PackedCSR.from_npy_cached(os.path.join(self.temp_dir, "sparse_nondiag_adj_nbr_1.npy"))
        self.node_embedding_init = M.MultipleEmbedding(embeddings, self.dimensions, False, self.num_list, targets)
        self.higashi_model = M.Hyper_SAGNN(n_head=8, d_model=self.dimensions, d_k=16, d_v=16, diag_mask=True,
                                           bottle_neck=self.dimensions, attribute_dict=attr, cell_feats=self.cell_feats,
                                           encoder_dynamic_nn=self.node_embedding_init, encoder_static_nn=self.node_embedding_init,
                                           chrom_num=C)
        if torch.cuda.is_available() and self.config.get("prefit_cell_autoencoder", True):
            self.prefit_cell_autoencoder(300)                # Higashi_wrapper.py:827-831
        per = int(self.batch_size / (self.neg_num + 1))
        self.training_data_generator = M.DataGenerator(train_data, train_chrom, train_weight, per, True, self.num_list, k=1)
        self.validation_data_generator = M.DataGenerator(test_data, test_chrom, test_weight, per, False, self.num_list, k=1)
        self.sampler = NegativeSampler(self.csr, self.num_list, self.max_bin)
        self.nbr = self.nbr_w = None

    def prefit_cell_autoencoder(self, iterations=300, lr=1e-3):
        """The 300-iteration pre-fit of the cell branch that prep_model runs right after building the embedding
        (Higashi_wrapper.py:827-831 -> TiedAutoEncoder.fit, Modules.py:290-330, with sparse=False: plain MSE against the targets,
        Adam lr 1e-3, early stop once 50 iterations are done and the loss has not improved by 1 % for 30 iterations):
          * PCA initialisation when the embedding is narrower than the input (weight_list[0] = reverse_weight_list[-1] =
            pca.components_, Modules.py:293-296);
          * the reconstruction loss and its gradients come from the training kernels (hsg_recon_loss / hsg_recon_apply: the cell
            auto-encoder over ALL cells -- the reference draws 1024 cells with replacement per iteration, an unseeded draw; with
            the whole table the loss is its expectation);
          * Adam runs on the flat parameter views of the auto-encoder only, as `self.parameters()` of the branch does.
        Needs a GPU (no CPU fallback); returns the list of losses."""
        if not torch.cuda.is_available():
            raise RuntimeError("prefit_cell_autoencoder runs on the B200 CUDA path only (no CPU fallback)")
        ae = self.node_embedding_init.wstack[0]
        x = np.asarray(self.feats[0], dtype=np.float32)
        if ae.shape_list[1] < x.shape[1]:
            from sklearn.decomposition import PCA
            comp = torch.from_numpy(PCA(n_components=ae.shape_list[1]).fit(x).components_.astype(np.float32))
            with torch.no_grad():
                ae.weight_list[0].copy_(comp)
                ae.reverse_weight_list[-1].copy_(comp)
        eng = self._engine(1)
        pfx = eng.embed_prefix
        names = ["%s.wstack.0.%s" % (pfx, n) for n in ("weight_list.0", "bias_list.0", "reverse_weight_list.0", "recon_bias_list.0")]
        params, grads = eng.named_params(), eng.named_grads()
        views = [params[n] for n in names if n in params]
        for v, n in zip(views, [n for n in names if n in params]):
            v.grad = None
        opt = torch.optim.Adam(views, lr=lr)
        losses, best, stale = [], None, 0
        for i in range(int(iterations)):
            eng.zero_grad()
            loss, _ = eng.recon_loss()
            eng.recon_apply(1.0)
            for v, n in zip(views, [n for n in names if n in params]):
                v.grad = grads[n]
            opt.step()
            losses.append(loss)
            if best is None:
                best = loss
            if i >= 50:                                      # Modules.py:317-326
                if loss < best * 0.99:
                    best, stale = loss, 0
                else:
                    stale += 1
                if stale >= 30:
                    break
        self._sync_back(eng)
        eng.close()
        return losses

    # ---- training --------------------------------------------------------------------------------------------------------
    def _engine(self, stage):
        sd = self.higashi_model.state_dict()
        cell_table = None
        if stage >= 2:
            cell_table = self.node_embedding_init.off_hook_embedding[0].embedding
        eng = TrainEngine(sd, self.num, self.feats, self.cell_feats, device=torch.cuda.current_device(), cell_table=cell_table,
                          cell_on_hook=(stage == 1))
        eng.set_options(cell2weight=self.cell_feats1, rank_thres=float(self.rank_thres))
        return eng

    def _batch(self, gen, pair_ratio, training):
        e, c, w, chroms = gen.next_iter()
        chrom = int(chroms[0])
        neg = self.sampler.sample(e, chrom, self.neg_num, pair_ratio) if self.neg_num > 0 else np.zeros((0, 3), dtype=np.int64)
        corr = 1.0 if self.mode == "classification" else 0.0
        x = np.concatenate([e, neg]).astype(np.int64)
        y = np.concatenate([np.ones(len(e)), np.zeros(len(neg))]).astype(np.float32)
        ww = np.concatenate([w.reshape(-1), np.full(len(neg), corr)]).astype(np.float32)
        perm = np.random.permutation(len(x))
        return x[perm], y[perm], ww[perm], chrom

    def train(self, stage, epochs, dynamic_pair_ratio=False, pair_ratio=0.5, beta=1e-2, save_embed=False, save_name=""):
        eng = self._engine(stage)
        # training batches are produced on the device (negatives, labels, neighbour lists: hsg_sample_batch); the host
        # restatement (NegativeSampler + batches.neighbor_csr) stays for validation batches and as the tests' cross-check
        eng.set_graph(self.csr, self.nbr, self.nbr_w)
        tr = parallel.DataParallelTrainer(eng, lr=1e-3)
        mode = LOSS_MODES[self.mode]
        best, no_improve = 1000.0, 0
        n_batches = 0
        for epoch_i in range(epochs):
            t0, tot = time.time(), 0.0
            for _ in range(self.update_num_per_training_epoch):
                e, _, w, chroms = self.training_data_generator.next_iter()
                chrom = int(chroms[0])
                n_batches += 1
                e = np.asarray(e, dtype=np.int64)
                e[:, 1:] = np.sort(e[:, 1:], axis=1)
                B, nnz = eng.sample_batch(e, chrom, self.neg_num, pair_ratio, self.max_bin, pos_w=np.asarray(w).reshape(-1), training=True,
                                          classification=(self.mode == "classification"), seed=(int(self.config.get("random_seed", 0)) << 20) + n_batches)
                tot += tr.step(None, chrom, None, None, None, stage=min(stage, 2), loss_mode=mode, train_mode=True, beta=beta,
                               median_sparsity=self.median_total_sparsity_cell, contractive_weight=self.contractive_loss_weight if stage == 1 else 0.0,
                               sampled=(B, nnz))
            bce = tot / self.update_num_per_training_epoch
            vt = 0.0
            eng.set_dropout(False)
            for _ in range(self.update_num_per_eval_epoch):
                x, y, w, chrom = self._batch(self.validation_data_generator, pair_ratio, False)
                nb = batches.neighbor_csr(self.csr, x, chrom, self.num_list, nbr=self.nbr, nbr_w=self.nbr_w) if stage >= 2 else None
                vt += eng.forward(x, chrom, nb, y=y, w=w, stage=min(stage, 2), loss_mode=mode)[1]
            print("[ Epoch %d of %d ] - (Train) bce: %7.4f  - (Valid) bce: %7.4f  elapse: %.3f s" %
                  (epoch_i, epochs, bce, vt / self.update_num_per_eval_epoch, time.time() - t0))
            if dynamic_pair_ratio:
                pair_ratio = min(0.5, pair_ratio + 0.1)
            else:
                if bce < best - 1e-3:
                    best, no_improve = bce, 0
                else:
                    no_improve += 1
                if no_improve >= 4:
                    break
            self._sync_back(eng)
            if save_embed:
                self.save_embeddings()
            if epoch_i % 5 == 0:
                torch.save({"model_link": self.higashi_model.state_dict(), "epoch": epoch_i}, self.save_path + save_name)
        self._sync_back(eng)
        eng.close()

    def _sync_back(self, eng):
        sd = {k: v.detach().cpu() for k, v in eng.named_params().items()}
        self.higashi_model.load_state_dict(sd, strict=False)

    def save_embeddings(self):
        """Higashi_wrapper.py:1218-1245: `embed/{name}_{i}_origin.npy` for the cell table (i=0) and every chromosome."""
        emb = self.node_embedding_init
        for i in range(len(emb.wstack)):
            t = emb.branch_table(i).numpy()
            if i == 0:
                self.cell_embeddings = t
            np.save(os.path.join(self.embed_dir, "%s_%d_origin.npy" % (self.embedding_name, i)), t)

    def _save_stage(self, n):
        torch.save({"model_link": self.higashi_model.state_dict()}, self.save_path + "_stage%d" % n)
        torch.save(self.higashi_model, self.save_path + "_stage%d_model" % n)

    def _load_stage(self, n):
        import higashi_b200
        higashi_b200.install_aliases()
        self.higashi_model = torch.load(self.save_path + "_stage%d_model" % n, map_location="cpu", weights_only=False)
        self.node_embedding_init = self.higashi_model.encode1.static_nn

    def train_for_embeddings(self, max_epochs=None):
        print("First stage training")
        self.train(1, self.embedding_epoch if max_epochs is None else max_epochs, dynamic_pair_ratio=True, pair_ratio=0.0, beta=1e-2,
                   save_embed=True, save_name="_stage1")
        self._save_stage(1)

    def train_for_imputation_nbr_0(self):
        self.train_for_imputation_no_nbr()

    def _filter_generators(self):
        mx, mn = self.config["maximum_distance"], self.config["minimum_distance"]
        max_bin = int(1e5) if mx < 0 else int(mx / self.res)
        min_bin = 0 if mn < 0 else int(mn / self.res)
        self.training_data_generator.filter_edges(min_bin, max_bin)
        self.validation_data_generator.filter_edges(min_bin, max_bin)

    def train_for_imputation_no_nbr(self):
        self._load_stage(1)
        self.save_embeddings()
        self.node_embedding_init.off_hook([0])
        self._filter_generators()
        self.higashi_model.encode1.dynamic_nn = M.GraphSageEncoder_with_weights(
            features=self.node_embedding_init, linear_features=self.node_embedding_init, feature_dim=self.dimensions,
            embed_dim=self.dimensions, num_sample=8, gcn=False, num_list=self.num_list, transfer_range=0,
            start_end_dict=self.start_end_dict, pass_pseudo_id=False, remove=True, pass_remove=False)
        print("Second stage training")
        self.nbr = self.nbr_w = None
        self.train(2, self.no_nbr_epoch, pair_ratio=0.5, beta=1e-3, save_name="_stage2")
        self._save_stage(2)

    def get_cell_neighbor(self, start=0):
        """Higashi_wrapper.py:1303-1324 on the device (hsg_build_neighbors); returns 1-based ids (self first) and weights."""
        from . import _lib
        ctx = _lib.Context(torch.cuda.current_device())
        ctx.set_dims(self.dimensions, int(self.num[0]), [int(v) for v in self.num[1:]], 1)
        nbr0, w = ctx.build_neighbors(np.ascontiguousarray(self.cell_embeddings, dtype=np.float32), self.neighbor_num)
        ctx.close()
        return nbr0.astype(np.int64) + 1, w

    def train_for_imputation_with_nbr(self):
        self._load_stage(2)
        self.save_embeddings()
        self._filter_generators()
        print("getting cell nbr's nbr list")
        self.nbr, self.nbr_w = self.get_cell_neighbor(0)
        cell_neighbor_list = np.array([[]] + [list(r) for r in self.nbr], dtype=object)
        weight_dict = {(int(c), i + 1): np.float32(w) for i in range(len(self.nbr)) for c, w in zip(self.nbr[i], self.nbr_w[i])}
        np.save(os.path.join(self.temp_dir, "weighted_info.npy"), np.array([cell_neighbor_list, weight_dict], dtype=object),
                allow_pickle=True)
        print("Final stage training")
        self.train(3, self.with_nbr_epoch, pair_ratio=0.5, beta=1e-3, save_name="_stage3")
        self._save_stage(3)

    # ---- imputation --------------------------------------------------------------------------------------------------------
    def _impute(self, stage, k, weighted_info):
        self._load_stage(stage)
        name = "%s_nbr_%d_impute" % (self.embedding_name, k)
        sparse_path = os.path.join(self.temp_dir, "sparse_nondiag_adj_nbr_1.npy")
        if self.gpu_num < 2:
            impute_process(self.config_path, self.higashi_model, name, self.mode, 0, int(self.num[0]), sparse_path, weighted_info)
        else:
            n_gpu = min(self.gpu_num, torch.cuda.device_count())
            parallel.impute_multi_gpu(self.config_path, self.save_path + "_stage%d_model" % stage, name, self.mode, int(self.num[0]),
                                      sparse_path, weighted_info, gpus=list(range(n_gpu)), impute_list=self.impute_list,
                                      temp_dir=self.temp_dir)

    def impute_no_nbr(self):
        self._impute(2, 0, None)

    def impute_with_nbr(self):
        self._impute(3, self.neighbor_num - 1, os.path.join(self.temp_dir, "weighted_info.npy"))

    # ---- readers -----------------------------------------------------------------------------------------------------------
    def fetch_cell_embeddings(self):
        if self.cell_embeddings is None:
            self.cell_embeddings = np.load(os.path.join(self.embed_dir, "%s_%d_origin.npy" % (self.embedding_name, 0)))
        return self.cell_embeddings

    def fetch_map(self, chrom, cell):
        """Higashi_wrapper.py:1610-1641: (raw, imputed without neighbours, imputed with neighbours) as scipy CSR."""
        from scipy.sparse import csr_matrix
        c = self.chrom_list.index(chrom)
        s, e = self.chrom_start_end[c]
        size = int(e - s)
        out = []
        for k in (0, self.neighbor_num - 1):
            try:
                f = sink.read_container(self.temp_dir, chrom, "%s_nbr_%d_impute" % (self.embedding_name, k))
                co = f["coordinates"].astype("int")
                m = csr_matrix((f["cell_%d" % cell], (co[:, 0], co[:, 1])), shape=(size, size), dtype="float32")
                out.append(m + m.T)
            except Exception:
                out.append(np.zeros((size, size)))
        return self.csr.matrix(c, cell), out[0], out[1]
