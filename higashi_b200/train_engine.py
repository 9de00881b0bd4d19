"""Host driver of the training-side kernels (hsg_score_fwd / hsg_score_bwd).

Parameters and gradients live in two flat CUDA tensors; `named_params()` / `named_grads()` hand out views under the
reference's state_dict names, so `torch.optim.Adam` can run on the views and ONE `torch.distributed.all_reduce`
on `self.grads` is the whole data-parallel exchange (Higashi_wrapper.py:1006-1010 is where it slots in)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib

LOSS_CLASSIFICATION, LOSS_REGRESSION, LOSS_ZINB, LOSS_RANK = 0, 1, 2, 3
LOSS_MODES = {"classification": 0, "regression": 1, "zinb": 2, "rank": 3}
_FROZEN = {"W_ATTR", "W_CELL_FEATS"}


def _np(t):
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t)


def neighbors_to_csr(to_neighs, n_rows: int, bin_base: int):
    """to_neighs_to_mask output (indices [2,nnz], values [nnz], unique_nodes [U]; Higashi_wrapper.py:171-221) ->
    CSR over the bin-token rows with chromosome-local columns."""
    ind, val, uniq = [_np(t) for t in to_neighs]
    rows = ind[0].astype(np.int64)
    if len(rows) > 1 and np.any(np.diff(rows) < 0):
        order = np.argsort(rows, kind="stable")
        rows, ind, val = rows[order], ind[:, order], val[order]
    indptr = np.zeros(n_rows + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=n_rows), out=indptr[1:])
    cols = (uniq[ind[1]] - bin_base).astype(np.int32)
    return indptr, np.ascontiguousarray(cols), np.ascontiguousarray(val, dtype=np.float32)


def grad_presence(names, chrom: int, stage: int, zinb: bool, cell_on_hook: bool):
    """Which parameter tensors receive a gradient for a batch of chromosome `chrom` (the others keep grad = None, so Adam skips
    them exactly as in the reference, where untouched branches have no .grad)."""
    pat = []
    for n in names:
        if ".wstack." in n:
            b = int(n.split(".wstack.")[1].split(".")[0])
            on = (b == chrom + 1) or (b == 0 and stage == 1 and cell_on_hook)
        else:
            unused_head = (n.startswith("pff_classifier_var") or n.startswith("pff_classifier_proba") or
                           n.startswith("extra_proba.") or n.startswith("extra_proba3"))
            on = not ((unused_head and not zinb) or (stage == 1 and n.startswith("encode1.dynamic_nn.nn")))
        pat.append(bool(on))
    return pat


class TrainEngine:
    def __init__(self, sd: Dict[str, object], num: Sequence[int], feats: list, cell_feats, device: int = 0,
                 cell_table=None, embed_prefix: str = "encode1.static_nn", cell_on_hook: bool = False):
        self.ctx = _lib.Context(device)
        self.device = torch.device("cuda", device)
        self.num = [int(v) for v in num]
        self.num_list = np.cumsum(self.num)
        self.d = int(_np(sd["layer_norm1.weight"]).shape[0])
        n_chrom = len(self.num) - 1
        self.ctx.set_dims(self.d, self.num[0], self.num[1:], int(_np(cell_feats).shape[-1]))
        self.ctx.set_weight("W_ATTR", _np(sd["attribute_dict_embedding.weight"]).astype(np.float32))
        self.ctx.set_weight("W_CELL_FEATS", _np(cell_feats).astype(np.float32))
        # frozen cell table (stage >= 2: off_hook([0]), Higashi_wrapper.py:1384)
        if cell_table is not None:
            self.ctx.set_table(0, _np(cell_table).astype(np.float32))
        else:
            self.ctx.embed_table(0, _np(feats[0]), _np(sd["%s.wstack.0.weight_list.0" % embed_prefix]),
                                 _np(sd["%s.wstack.0.bias_list.0" % embed_prefix]))
        # flat layout: weight slots in enum order, then per-branch projections
        self.layout = {}
        off = 0
        off_slot = np.full(len(_lib.WEIGHT_SLOTS), -1, dtype=np.int64)
        key_of_slot = {v: k for k, v in _lib.STATE_KEYS.items()}
        for i, slot in enumerate(_lib.WEIGHT_SLOTS):
            key = key_of_slot.get(slot)
            if slot in _FROZEN or key is None or key not in sd:
                continue
            shape = tuple(_np(sd[key]).shape)
            n = int(np.prod(shape))
            self.layout[key] = (off, shape)
            off_slot[i] = off
            off += (n + 3) & ~3
        off_pw = np.full(n_chrom + 1, -1, dtype=np.int64)
        off_pb = np.full(n_chrom + 1, -1, dtype=np.int64)
        for b in range(n_chrom + 1):
            for which, arr in (("weight_list", off_pw), ("bias_list", off_pb)):
                key = "%s.wstack.%d.%s.0" % (embed_prefix, b, which)
                shape = tuple(_np(sd[key]).shape)
                self.layout[key] = (off, shape)
                arr[b] = off
                off += (int(np.prod(shape)) + 3) & ~3
        self.cell_on_hook = cell_on_hook
        if cell_on_hook:                       # stage 1: decoder half of the cell auto-encoder (reconstruction loss)
            for which in ("reverse_weight_list", "recon_bias_list"):
                key = "%s.wstack.0.%s.0" % (embed_prefix, which)
                shape = tuple(_np(sd[key]).shape)
                self.layout[key] = (off, shape)
                off += (int(np.prod(shape)) + 3) & ~3
        self.total = off
        self.params = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        # gradient buffer + one presence flag per parameter tensor behind it (one storage: one all_reduce moves both)
        self._gbuf = torch.zeros(self.total + len(self.layout), dtype=torch.float32, device=self.device)
        self.grads = self._gbuf[:self.total]
        self.presence = self._gbuf[self.total:]
        for key, (o, shape) in self.layout.items():
            n = int(np.prod(shape))
            self.params[o:o + n] = torch.from_numpy(_np(sd[key]).astype(np.float32).reshape(-1)).to(self.device)
        for b in range(0 if cell_on_hook else 1, n_chrom + 1):
            f = np.ascontiguousarray(_np(feats[b]), dtype=np.float32)
            self.ctx._check(self.ctx.lib.hsg_train_set_features(self.ctx.h, b, f.ctypes.data_as(C.POINTER(C.c_float)), f.shape[0], f.shape[1]))
        i64p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        self._offs = (off_slot, off_pw, off_pb)
        self.ctx._check(self.ctx.lib.hsg_train_bind(self.ctx.h, C.c_void_p(self.params.data_ptr()), C.c_void_p(self.grads.data_ptr()),
                                                    self.total, i64p(off_slot), i64p(off_pw), i64p(off_pb)))

        if cell_on_hook:
            self.ctx._check(self.ctx.lib.hsg_train_bind_recon(
                self.ctx.h, self.layout["%s.wstack.0.reverse_weight_list.0" % embed_prefix][0],
                self.layout["%s.wstack.0.recon_bias_list.0" % embed_prefix][0]))
        self.embed_prefix = embed_prefix

    def close(self):
        self.ctx.close()

    def recon_loss(self):
        """(mse over all cells, ||grad wstack.0.weight_list.0||) of the cell auto-encoder; gradients stay in a scratch block."""
        loss, gn = C.c_float(0), C.c_float(0)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx._check(self.ctx.lib.hsg_recon_loss(self.ctx.h, C.byref(loss), C.byref(gn), C.c_void_p(stream)))
        return float(loss.value), float(gn.value)

    def recon_apply(self, scale: float):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx._check(self.ctx.lib.hsg_recon_apply(self.ctx.h, C.c_float(scale), C.c_void_p(stream)))

    def named_params(self):
        return {k: self.params[o:o + int(np.prod(s))].view(s) for k, (o, s) in self.layout.items()}

    def named_grads(self):
        return {k: self.grads[o:o + int(np.prod(s))].view(s) for k, (o, s) in self.layout.items()}

    def zero_grad(self):
        self.grads.zero_()

    # ---- device batch producer (generate_negative_cpu + one_thread_generate_neg, Higashi_wrapper.py:95-333) ------------------
    def set_graph(self, csr, nbr=None, nbr_w=None):
        """Contacts (PackedCSR) and optional kNN cell graph (1-based ids, self first) the batch producer samples against."""
        self.ctx.set_contacts(csr.row_ptr, csr.col, csr.val)
        if nbr is not None:
            self.ctx.set_neighbors(np.ascontiguousarray(np.asarray(nbr) - 1, dtype=np.int32), np.ascontiguousarray(nbr_w, dtype=np.float32))

    def sample_batch(self, pos, chrom: int, neg_num: int, pair_ratio: float, max_bin: int, pos_w=None, training: bool = True,
                     classification: bool = True, wide_check: bool = True, seed: int = 0):
        """Builds [positives ; negatives], labels, weights and the neighbour CSR on the device; returns (B, nnz) for
        forward(None, chrom, sampled=(B, nnz))."""
        pos = np.ascontiguousarray(_np(pos), dtype=np.int64)
        pw = None if pos_w is None else np.ascontiguousarray(_np(pos_w), dtype=np.float32).reshape(-1)
        B, nnz = C.c_int32(0), C.c_int64(0)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx._check(self.ctx.lib.hsg_sample_batch(
            self.ctx.h, pos.ctypes.data_as(C.POINTER(C.c_int64)), None if pw is None else pw.ctypes.data_as(C.POINTER(C.c_float)),
            pos.shape[0], int(chrom), int(neg_num), C.c_float(pair_ratio), int(max_bin), 1 if wide_check else 0, 1 if training else 0,
            C.c_float(1.0 if classification else 0.0), C.c_uint64(int(seed) & (2 ** 64 - 1)), C.byref(B), C.byref(nnz), C.c_void_p(stream)))
        return int(B.value), int(nnz.value)

    def set_removal_draws(self, draws=None):
        """Reproducibility hook (hsg_sample_set_removal_draws): uniform draws of the partner removal, one per bin-token row."""
        if draws is None:
            self.ctx._check(self.ctx.lib.hsg_sample_set_removal_draws(self.ctx.h, None, 0))
            return
        d = np.ascontiguousarray(_np(draws), dtype=np.float32).reshape(-1)
        self.ctx._check(self.ctx.lib.hsg_sample_set_removal_draws(self.ctx.h, d.ctypes.data_as(C.POINTER(C.c_float)), d.size))

    def sample_batch_begin(self, pos, chrom: int, neg_num: int, pair_ratio: float, max_bin: int, pos_w=None, training: bool = True,
                           classification: bool = True, wide_check: bool = True, seed: int = 0):
        """Asynchronous first half of sample_batch (own stream): call it for batch i+1 before stepping on batch i."""
        pos = np.ascontiguousarray(_np(pos), dtype=np.int64)
        pw = None if pos_w is None else np.ascontiguousarray(_np(pos_w), dtype=np.float32).reshape(-1)
        self.ctx._check(self.ctx.lib.hsg_sample_batch_begin(
            self.ctx.h, pos.ctypes.data_as(C.POINTER(C.c_int64)), None if pw is None else pw.ctypes.data_as(C.POINTER(C.c_float)),
            pos.shape[0], int(chrom), int(neg_num), C.c_float(pair_ratio), int(max_bin), 1 if wide_check else 0, 1 if training else 0,
            C.c_float(1.0 if classification else 0.0), C.c_uint64(int(seed) & (2 ** 64 - 1))))

    def sample_batch_end(self):
        """Second half: returns (B, nnz); the batch is then in the training workspace."""
        B, nnz = C.c_int32(0), C.c_int64(0)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx._check(self.ctx.lib.hsg_sample_batch_end(self.ctx.h, C.byref(B), C.byref(nnz), C.c_void_p(stream)))
        return int(B.value), int(nnz.value)

    def fetch_sampled(self, B: int, nnz: int):
        """(x, y, w, indptr, rowcnt, cols, vals) of the batch the producer left on the device (tests / debugging)."""
        x = np.empty((B, 3), np.int64); y = np.empty(B, np.float32); w = np.empty(B, np.float32)
        indptr = np.empty(2 * B + 1, np.int32); rowcnt = np.empty(2 * B, np.int32)
        cols = np.empty(max(nnz, 1), np.int32); vals = np.empty(max(nnz, 1), np.float32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        pt = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        self.ctx._check(self.ctx.lib.hsg_sampled_fetch(self.ctx.h, pt(x, C.c_int64), pt(y, C.c_float), pt(w, C.c_float), pt(indptr, C.c_int32),
                                                       pt(rowcnt, C.c_int32), pt(cols, C.c_int32), pt(vals, C.c_float), C.c_void_p(stream)))
        return x, y, w, indptr, rowcnt, cols[:nnz], vals[:nnz]

    def forward(self, x, chrom: int, to_neighs=None, y=None, w=None, stage: int = 2, loss_mode: int = LOSS_CLASSIFICATION,
                only_model: bool = False, want_outputs: bool = True, want_mean: bool = True, sampled=None):
        """One minibatch of ONE chromosome (DataGenerator.next_iter draws one chromosome per batch, Modules.py:1205).
        Returns (mean float32 [B] or None, loss float or None).  x = None with sampled = (B, nnz): the batch sample_batch left
        on the device."""
        if x is None:
            B, nnz = sampled
            mean = np.empty(B, dtype=np.float32) if (want_outputs and want_mean) else None
            loss = np.zeros(1, dtype=np.float32) if want_outputs else None
            stream = torch.cuda.current_stream(self.device).cuda_stream
            fp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))
            self.ctx._check(self.ctx.lib.hsg_score_fwd(self.ctx.h, int(stage), int(chrom), None, B, None, None, None, int(nnz), None, None,
                                                       int(loss_mode), 1 if only_model else 0, fp(mean), fp(loss), C.c_void_p(stream)))
            return mean, (None if loss is None else float(loss[0]))
        x = np.ascontiguousarray(_np(x), dtype=np.int64)
        B = x.shape[0]
        bin_base = int(self.num_list[chrom]) + 1
        p = lambda a, t: None if a is None else a.ctypes.data_as(C.POINTER(t))
        indptr = cols = vals = None
        nnz = 0
        if stage == 2:
            if isinstance(to_neighs, tuple) and len(to_neighs) == 3 and _np(to_neighs[0]).ndim == 1:
                indptr, cols, vals = to_neighs                       # already CSR (batches.neighbor_csr)
            else:
                indptr, cols, vals = neighbors_to_csr(to_neighs, 2 * B, bin_base)
            nnz = int(len(cols))
        yv = None if y is None else np.ascontiguousarray(_np(y), dtype=np.float32).reshape(-1)
        wv = None if w is None else np.ascontiguousarray(_np(w), dtype=np.float32).reshape(-1)
        mean = np.empty(B, dtype=np.float32) if (want_outputs and want_mean) else None
        loss = np.zeros(1, dtype=np.float32) if (want_outputs and y is not None) else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx._check(self.ctx.lib.hsg_score_fwd(self.ctx.h, int(stage), int(chrom), p(x, C.c_int64), B, p(indptr, C.c_int32),
                                                   p(cols, C.c_int32), p(vals, C.c_float), nnz, p(yv, C.c_float), p(wv, C.c_float),
                                                   int(loss_mode), 1 if only_model else 0, p(mean, C.c_float), p(loss, C.c_float),
                                                   C.c_void_p(stream)))
        return mean, (None if loss is None else float(loss[0]))

    def set_options(self, cell2weight=None, rank_thres: float = 0.0):
        """cell2weight (zinb mode) / rank threshold (rank mode): Higashi_wrapper.py:898, 884."""
        cw = None if cell2weight is None else np.ascontiguousarray(_np(cell2weight), dtype=np.float32)
        self.ctx._check(self.ctx.lib.hsg_train_options(self.ctx.h, None if cw is None else cw.ctypes.data_as(C.POINTER(C.c_float)),
                                                       C.c_float(rank_thres)))

    def set_dropout(self, enabled: bool, seed: int = 0):
        """model.train() / model.eval() for the dropout layers (Modules.py:685-686)."""
        self.ctx._check(self.ctx.lib.hsg_train_set_dropout(self.ctx.h, 1 if enabled else 0, C.c_uint64(int(seed) & (2 ** 64 - 1))))

    def heads(self, B: int, zinb: bool = False):
        """(mean, var, proba, pred) of the last forward batch; var / proba / pred only after a zinb forward."""
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        mean = np.empty(B, np.float32)
        var = np.empty(B, np.float32) if zinb else None
        proba = np.empty(B, np.float32) if zinb else None
        pred = np.empty(B, np.float32) if zinb else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx._check(self.ctx.lib.hsg_score_heads(self.ctx.h, fp(mean), None if var is None else fp(var),
                                                     None if proba is None else fp(proba), None if pred is None else fp(pred),
                                                     C.c_void_p(stream)))
        return mean, var, proba, pred

    def heads_device(self, B: int, zinb: bool = False):
        """(mean, var, proba) of the last forward batch as CUDA tensors [B] (no host round trip: hsg_score_heads copies with
        cudaMemcpyDefault, so device destinations are device-to-device copies on the caller's stream)."""
        n = 3 if zinb else 1
        out = torch.empty((n, B), dtype=torch.float32, device=self.device)
        ptr = lambda i: C.cast(C.c_void_p(out[i].data_ptr()), C.POINTER(C.c_float))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx._check(self.ctx.lib.hsg_score_heads(self.ctx.h, ptr(0), ptr(1) if zinb else None, ptr(2) if zinb else None, None,
                                                     C.c_void_p(stream)))
        return (out[0], out[1], out[2]) if zinb else (out[0], None, None)

    def set_head_grads(self, dmean, dvar=None, dproba=None):
        """Upstream gradients (CUDA float32 tensors [B]) of the heads of the last forward, from an external loss."""
        keep = [None if t is None else t.detach().to(self.device, torch.float32).contiguous().reshape(-1) for t in (dmean, dvar, dproba)]
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx._check(self.ctx.lib.hsg_score_set_head_grads(self.ctx.h, ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), C.c_void_p(stream)))
        torch.cuda.current_stream(self.device).synchronize()        # `keep` may be freed once the copies are done

    def backward(self):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx._check(self.ctx.lib.hsg_score_bwd(self.ctx.h, C.c_void_p(stream)))

    def allreduce_grads(self):
        """The ONE collective of data-parallel training: sum over ranks, then mean (Adam then steps identically)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM)
            self.grads.div_(dist.get_world_size())
