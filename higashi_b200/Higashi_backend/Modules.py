"""Host-side mirror of the reference's `higashi/Higashi_backend/Modules.py` for the scorer hot path.

Same class names, constructor signatures, parameter names and method names as the reference, so that
reference `state_dict`s load unchanged and scripts written against the reference keep working.  The classes
here are parameter containers plus marshalling: every tensor computation of the hot path runs in libhsg
(CUDA, sm_100a) through `higashi_b200.engine.ScorerEngine`; there is no eager/CPU implementation and the calls
raise if the library or a GPU is missing.

reference symbol                               -> here
Hyper_SAGNN                (Modules.py:657)    -> Hyper_SAGNN      (.predict / .forward in imputation mode)
EncoderLayer               (:1098)             -> EncoderLayer     (container)
MultiHeadAttention         (:975)              -> MultiHeadAttention (container)
PositionwiseFeedForward    (:834), FeedForward (:896), ScaledDotProductAttention (:929) -> containers
MultipleEmbedding          (:501)              -> MultipleEmbedding (.off_hook/.on_hook/.forward on tables)
SparseEmbedding            (:101), TiedAutoEncoder (:146) -> containers (+ table gather)
GraphSageEncoder_with_weights (:1387)          -> same (.start_fix/.fix_cell2/.forward_off_hook/.forward_on_hook)
MeanAggregator[_with_weights] (:1237,:1305)    -> containers
DataGenerator              (:1165)             -> DataGenerator (host sampler, numpy)
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import engine as _engine
from .Functions import swish  # noqa: F401  (re-exported like the reference's star import)

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
activation_func = swish


def _require_cuda(what: str):
    if not torch.cuda.is_available():
        raise RuntimeError("%s runs on the B200 CUDA path only (libhsg); no CPU fallback exists" % what)


# ----------------------------------------------------------------------------- small containers
class SparseEmbedding(nn.Module):
    """Row table `embedding[x, :]` (Modules.py:101-142).  Holds the raw node features of one node type, or an
    off-hooked [n, d] table."""

    def __init__(self, embedding_weight, sparse=False, cpu=False):
        super().__init__()
        self.sparse = sparse
        self.cpu_flag = cpu
        if isinstance(embedding_weight, np.ndarray):
            embedding_weight = torch.from_numpy(np.asarray(embedding_weight))
        self.embedding = embedding_weight

    def forward(self, x):
        return self.embedding[x.to(self.embedding.device), :]


class TiedAutoEncoder(nn.Module):
    """Parameters of the (un)tied auto-encoder branch (Modules.py:146-233): weight_list / reverse_weight_list /
    bias_list / recon_bias_list.  Only the encoder is on the hot path and it runs as `hsg_embed_table`."""

    def __init__(self, shape_list, use_bias=True, tied_list=None, add_activation=False, dropout=None,
                 layer_norm=False, activation=None):
        super().__init__()
        tied_list = [] if tied_list is None else tied_list
        self.add_activation = add_activation
        self.use_bias = use_bias
        self.shape_list = shape_list
        self.tied_list = tied_list
        w, rw, b, rb = [], [], [], []
        for i in range(len(shape_list) - 1):
            p = nn.Parameter(torch.empty(shape_list[i + 1], shape_list[i]))
            w.append(p)
            rw.append(p if i in tied_list else nn.Parameter(torch.empty(shape_list[i + 1], shape_list[i])))
            b.append(nn.Parameter(torch.empty(shape_list[i + 1])))
            rb.append(nn.Parameter(torch.empty(shape_list[i])))
        self.weight_list = nn.ParameterList(w)
        self.reverse_weight_list = nn.ParameterList(rw[::-1])
        self.bias_list = nn.ParameterList(b)
        self.recon_bias_list = nn.ParameterList(rb[::-1])
        self.layer_norm = nn.LayerNorm(shape_list[-1]) if layer_norm else None
        self.dropout = nn.Dropout(dropout) if dropout is not None else None
        self.input_dropout = nn.Dropout(0.1)
        self.reset_parameters()

    def reset_parameters(self):
        for wt in self.weight_list:
            nn.init.kaiming_uniform_(wt, a=0.0, mode="fan_in", nonlinearity="leaky_relu")
        for wt in self.reverse_weight_list:
            nn.init.kaiming_uniform_(wt, a=0.0, mode="fan_out", nonlinearity="leaky_relu")
        for i, bb in enumerate(self.bias_list):
            nn.init.uniform_(bb, -1 / math.sqrt(self.weight_list[i].shape[1]), 1 / math.sqrt(self.weight_list[i].shape[1]))
        for bb in self.recon_bias_list:
            nn.init.uniform_(bb, -1 / math.sqrt(max(1, bb.shape[0])), 1 / math.sqrt(max(1, bb.shape[0])))


class FeedForward(nn.Module):
    """Linear stack container (Modules.py:896-926): `w_stack.{i}.{weight,bias}`."""

    def __init__(self, dims, dropout=None, reshape=False, use_bias=True):
        super().__init__()
        self.w_stack = nn.ModuleList([nn.Linear(dims[i], dims[i + 1], use_bias) for i in range(len(dims) - 1)])
        self.dropout = nn.Dropout(dropout) if dropout is not None else None
        self.reshape = reshape


class PositionwiseFeedForward(nn.Module):
    """1x1-conv stack container (Modules.py:834-889): `w_stack.{i}.{weight[o,i,1],bias}`, `layer_norm.*`, `alpha`."""

    def __init__(self, dims, dropout=None, reshape=False, use_bias=True, residual=False, layer_norm=False):
        super().__init__()
        self.dims = dims
        self.w_stack = nn.ModuleList([nn.Conv1d(dims[i], dims[i + 1], 1, bias=use_bias) for i in range(len(dims) - 1)])
        self.reshape = reshape
        self.layer_norm = nn.LayerNorm(dims[0])
        self.dropout = nn.Dropout(dropout) if dropout is not None else None
        self.residual = residual
        self.layer_norm_flag = layer_norm
        self.alpha = nn.Parameter(torch.zeros(1))


class ScaledDotProductAttention(nn.Module):
    def __init__(self, temperature):
        super().__init__()
        self.temperature = temperature


class MultiHeadAttention(nn.Module):
    """Container for w_qs / w_ks / w_vs / fc1 / fc2 / layer_norm{1,2,3} / alpha_* (Modules.py:975-1027)."""

    def __init__(self, n_head, d_model, d_k, d_v, dropout, diag_mask, input_dim):
        super().__init__()
        self.d_model, self.input_dim, self.n_head, self.d_k, self.d_v = d_model, input_dim, n_head, d_k, d_v
        self.w_qs = nn.Linear(input_dim, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(input_dim, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(input_dim, n_head * d_v, bias=False)
        nn.init.normal_(self.w_qs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_ks.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_vs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_v)))
        self.attention = ScaledDotProductAttention(temperature=np.power(d_k, 0.5))
        self.fc1 = FeedForward([n_head * d_v, d_model], use_bias=False)
        self.fc2 = FeedForward([n_head * d_v, d_model], use_bias=False)
        self.layer_norm1 = nn.LayerNorm(input_dim)
        self.layer_norm2 = nn.LayerNorm(input_dim)
        self.layer_norm3 = nn.LayerNorm(input_dim)
        self.dropout = nn.Dropout(dropout) if dropout is not None else None
        self.diag_mask_flag = diag_mask
        self.diag_mask = None
        self.alpha_static = nn.Parameter(torch.zeros(1))
        self.alpha_dynamic = nn.Parameter(torch.zeros(1))


class EncoderLayer(nn.Module):
    """mul_head_attn + pff_n1 (+ the never-called pff_n2) + the two node-embedding modules (Modules.py:1098-1134)."""

    def __init__(self, n_head, d_model, d_k, d_v, dropout_mul, dropout_pff, diag_mask, bottle_neck,
                 dynamic_nn=None, static_nn=None):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.mul_head_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout_mul, diag_mask=diag_mask,
                                                input_dim=bottle_neck)
        self.pff_n1 = PositionwiseFeedForward([d_model, d_model, d_model], dropout=dropout_pff, residual=True, layer_norm=True)
        self.pff_n2 = PositionwiseFeedForward([bottle_neck, d_model, d_model], dropout=dropout_pff,
                                              residual=(bottle_neck == d_model), layer_norm=True)
        self.dynamic_nn = dynamic_nn
        self.static_nn = static_nn
        self.dropout = nn.Dropout(0.2)


# ----------------------------------------------------------------------------- node embedding
class MultipleEmbedding(nn.Module):
    """Routes node ids to per-type branches (cells, then one per chromosome) (Modules.py:501-654).

    `off_hook()` replaces a branch by its [n, d] table; the table is computed on the GPU by `hsg_embed_table`
    (TiedAutoEncoder.encoder over all nodes of the branch)."""

    def __init__(self, embedding_weights, dim, sparse=True, num_list=None, target_weights=None):
        super().__init__()
        if target_weights is None:
            target_weights = embedding_weights
        self.dim = dim
        self.num_list = torch.tensor([0] + [int(v) for v in num_list])
        self.embeddings = [SparseEmbedding(w, sparse) for w in embedding_weights]
        self.targets = [SparseEmbedding(w, sparse) for w in target_weights]
        self.input_size = [int(e.embedding.shape[-1]) for e in self.embeddings]
        self.layer_norm = nn.LayerNorm(self.dim)
        stack = [TiedAutoEncoder([self.input_size[0], self.dim], add_activation=False, tied_list=[])]
        for i in range(1, len(self.embeddings)):
            stack.append(TiedAutoEncoder([self.input_size[i], self.dim], add_activation=True, tied_list=[]))
        self.wstack = nn.ModuleList(stack)
        # registered under the same name as the reference so that `on_hook_embedding.{i}.1.*` state_dict keys
        # (aliases of wstack.{i}.*, Modules.py:562-564) load unchanged
        self.on_hook_embedding = nn.ModuleList([nn.Sequential(e, self.wstack[i]) for i, e in enumerate(self.embeddings)])
        self.on_hook_set = set(range(len(self.embeddings)))
        self.off_hook_embedding: List[Optional[SparseEmbedding]] = [None] * len(self.embeddings)
        self.features = self.forward

    # -- tables -------------------------------------------------------------
    def branch_table(self, index: int, ctx=None) -> torch.Tensor:
        """[n_index, d] table of one branch via the CUDA encoder kernel."""
        _require_cuda("MultipleEmbedding.off_hook")
        ae = self.wstack[index]
        own = ctx is None
        if own:
            from .. import _lib
            ctx = _lib.Context(torch.cuda.current_device())
            n = [int(self.num_list[i + 1] - self.num_list[i]) for i in range(len(self.wstack))]
            ctx.set_dims(self.dim, n[0], n[1:] if len(n) > 1 else [1], 1)
        feats = self.embeddings[index].embedding.detach().cpu().numpy().astype(np.float32)
        t = ctx.embed_table(index, feats, ae.weight_list[0].detach().cpu().numpy(), ae.bias_list[0].detach().cpu().numpy(),
                            want_copy=True)
        if own:
            ctx.close()
        return torch.from_numpy(t)

    def off_hook(self, off_hook_list=[]):
        idx = list(off_hook_list) if len(off_hook_list) else list(range(len(self.wstack)))
        for index in idx:
            ae = self.wstack[index]
            for plist in (ae.weight_list, ae.reverse_weight_list, ae.bias_list, ae.recon_bias_list):
                for prm in plist:
                    prm.requires_grad = False
            self.off_hook_embedding[index] = SparseEmbedding(self.branch_table(index), False)
            self.on_hook_set.discard(index)

    def on_hook(self, on_hook_list=[]):
        idx = list(on_hook_list) if len(on_hook_list) else list(range(len(self.wstack)))
        for index in idx:
            ae = self.wstack[index]
            for plist in (ae.weight_list, ae.reverse_weight_list, ae.bias_list, ae.recon_bias_list):
                for prm in plist:
                    prm.requires_grad = True
            self.on_hook_set.add(index)

    def forward(self, x, *args, route_nn=None):
        """Embedding lookup for off-hooked branches (table gather).  On-hook branches are trained by the CUDA
        training kernels and are not evaluated here."""
        shape = x.shape
        flat = x.reshape(-1)
        out = torch.zeros((flat.numel(), self.dim), dtype=torch.float32)
        nl = self.num_list
        for i in range(len(self.wstack)):
            m = (flat > nl[i]) & (flat <= nl[i + 1])
            if not bool(m.any()):
                continue
            if i in self.on_hook_set or self.off_hook_embedding[i] is None:
                raise RuntimeError("branch %d is on-hook: call off_hook() first (imputation) -- on-hook evaluation "
                                   "belongs to the training kernels" % i)
            out[m] = self.off_hook_embedding[i].embedding[(flat[m] - nl[i] - 1).cpu()]
        return out.view(*shape, self.dim)

    def start_fix(self):
        return

    def fix_cell(self, cell=None, bin_id=None):
        return


class MeanAggregator(nn.Module):
    def __init__(self, features, gcn=False, num_list=None, start_end_dict=None, pass_pseudo_id=False):
        super().__init__()
        self.features, self.gcn = features, gcn
        self.num_list = torch.as_tensor(np.asarray(num_list))
        self.start_end_dict, self.pass_pseudo_id = start_end_dict, pass_pseudo_id


class MeanAggregator_with_weights(nn.Module):
    def __init__(self, features, gcn=False, num_list=None, start_end_dict=None, pass_pseudo_id=False, remove=False,
                 pass_remove=False):
        super().__init__()
        self.features, self.gcn = features, gcn
        self.num_list = torch.as_tensor(np.asarray(num_list))
        self.start_end_dict, self.pass_pseudo_id = start_end_dict, pass_pseudo_id
        self.remove, self.pass_remove = remove, pass_remove


class GraphSageEncoder_with_weights(nn.Module):
    """GraphSAGE neighbour encoder (Modules.py:1387-1553).  Owns `nn` (Linear 2d -> d).  In imputation mode the
    per-cell work of `fix_cell2` (A_hat . E, concat, later swish(nn(.)) and the attention projections) is the
    K1a/K1b kernel pair; the object only records which cell is fixed."""

    def __init__(self, features, linear_features=None, feature_dim=64, embed_dim=64, num_sample=10, gcn=False,
                 num_list=None, transfer_range=0, start_end_dict=None, pass_pseudo_id=False, remove=False,
                 pass_remove=False):
        super().__init__()
        if gcn or transfer_range > 0:
            raise NotImplementedError("gcn=True / transfer_range>0 are never reached from the reference wrapper "
                                      "(Higashi_wrapper.py:1415-1422) and are out of scope")
        self.features, self.linear_features = features, linear_features
        self.feat_dim, self.embed_dim = feature_dim, embed_dim
        self.pass_pseudo_id = pass_pseudo_id
        self.aggregator = MeanAggregator_with_weights(features, gcn, num_list, start_end_dict, pass_pseudo_id, remove, pass_remove)
        self.linear_aggregator = MeanAggregator(linear_features, gcn, num_list, start_end_dict, pass_pseudo_id)
        self.num_sample, self.transfer_range, self.gcn = num_sample, transfer_range, gcn
        self.start_end_dict = start_end_dict
        self.nn = nn.Linear(2 * feature_dim, embed_dim)
        self.num_list = torch.as_tensor(np.asarray(num_list))
        self.fix = False
        self.fixed_cell = None
        self.cell_feats = None
        self.forward = self.forward_on_hook

    def start_fix(self):
        """Modules.py:1426-1429: cache the cell table (here: make sure the cell branch is off-hooked)."""
        self.fix = True
        if 0 in self.features.on_hook_set:
            self.features.off_hook([0])
        self.cell_feats = self.features.off_hook_embedding[0].embedding

    def fix_cell2(self, cell, bin_ids=None, sparse_matrix=None, local_transfer_range=0, route_nn_list=None):
        """Modules.py:1432-1469.  The adjacency comes from the engine's device-resident CSR (the `sparse_matrix`
        argument of the reference is what `prep_one` would have produced from the same data)."""
        self.fix = True
        self.fixed_cell = int(cell)
        self.local_transfer_range = int(local_transfer_range)      # Hyper_SAGNN.predict prepares the engine with it

    def forward_off_hook(self, nodes, to_neighs, *args, route_nn=None):
        raise RuntimeError("forward_off_hook is fused into Hyper_SAGNN.predict on the CUDA path")

    def forward_on_hook(self, nodes, to_neighs, *args, route_nn=None):
        raise RuntimeError("on-hook (training) forward belongs to the training kernels (DESIGN.md section 9)")


# ----------------------------------------------------------------------------- the scorer
class _HyperedgeScoreFn(torch.autograd.Function):
    """Hyper_SAGNN.forward as one autograd node: forward = hsg_score_fwd (no loss), backward = hsg_score_set_head_grads +
    hsg_score_bwd; parameter gradients are written to the module's parameters as a side effect of backward."""

    @staticmethod
    def forward(ctx, anchor, model, x, chrom, to_neighs, stage):
        eng = model._train_engine(stage)
        model._push_params(eng)
        model._tseed += 1
        eng.set_dropout(bool(model.training), seed=0x51ED0000 + model._tseed)
        xx = np.ascontiguousarray(np.asarray(x.detach().cpu() if hasattr(x, "detach") else x), dtype=np.int64)
        eng.forward(xx, chrom, to_neighs, stage=stage, loss_mode=2, only_model=bool(model.only_model), want_outputs=False)
        mean, var, proba = eng.heads_device(len(xx), zinb=True)          # stays on the device: no numpy round trip
        ctx.set_materialize_grads(False)
        ctx.model, ctx.eng, ctx.chrom, ctx.stage = model, eng, chrom, stage
        return mean.view(-1, 1), var.view(-1, 1), proba.view(-1, 1)

    @staticmethod
    def backward(ctx, gmean, gvar, gproba):
        eng, model = ctx.eng, ctx.model
        if gmean is None:
            gmean = torch.zeros_like(gvar if gvar is not None else gproba)
        eng.zero_grad()
        zinb = gvar is not None or gproba is not None
        eng.set_head_grads(gmean, gvar, gproba)
        eng.backward()
        torch.cuda.current_stream(eng.device).synchronize()
        model._pull_grads(eng, ctx.chrom, ctx.stage, zinb)
        return None, None, None, None, None, None


class Hyper_SAGNN(nn.Module):
    """Reference: Modules.py:657-827.  Same constructor and attributes; `predict` runs the fused CUDA scorer."""

    def __init__(self, n_head, d_model, d_k, d_v, diag_mask, bottle_neck, attribute_dict=None, cell_feats=None,
                 encoder_dynamic_nn=None, encoder_static_nn=None, chrom_num=1):
        super().__init__()
        if n_head != 8 or d_k != 16 or d_v != 16 or not diag_mask or bottle_neck != d_model:
            raise NotImplementedError("the CUDA scorer implements the configuration the reference wrapper builds: "
                                      "n_head=8, d_k=d_v=16, diag_mask=True, bottle_neck=d_model (Higashi_wrapper.py:834-845)")
        self.pff_classifier = PositionwiseFeedForward([d_model, d_model // 2, 1])
        self.pff_classifier_var = PositionwiseFeedForward([d_model, d_model // 2, 1])
        self.pff_classifier_proba = PositionwiseFeedForward([d_model, d_model // 2, 1])
        self.encode_list = []
        self.encode1 = EncoderLayer(n_head, d_model, d_k, d_v, dropout_mul=0.3, dropout_pff=0.4, diag_mask=diag_mask,
                                    bottle_neck=bottle_neck, dynamic_nn=encoder_dynamic_nn, static_nn=encoder_static_nn)
        self.diag_mask_flag = diag_mask
        self.layer_norm1 = nn.LayerNorm(d_model)
        self.layer_norm2 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(0.3)
        self.attribute_dict = None
        if attribute_dict is not None:
            self.attribute_dict = torch.from_numpy(np.asarray(attribute_dict, dtype=np.float32))
            input_size = self.attribute_dict.shape[-1] * 2 + cell_feats.shape[-1]
            self.extra_proba = FeedForward([input_size, 4, 1])
            self.extra_proba2 = FeedForward([input_size, 4, 1])
            self.extra_proba3 = FeedForward([input_size, 4, 1])
            self.attribute_dict_embedding = nn.Embedding(len(self.attribute_dict), 1, padding_idx=0)
            self.attribute_dict_embedding.weight = nn.Parameter(self.attribute_dict.clone(), requires_grad=False)
            self.cell_feats = torch.from_numpy(np.asarray(cell_feats, dtype=np.float32))
        self.only_distance = False
        self.only_model = False
        self.chrom_num = chrom_num
        self.d_model = d_model
        self._engine = None

    # -- engine plumbing ------------------------------------------------------
    def build_engine(self, csr=None, nbr=None, nbr_w=None, device_index: Optional[int] = None):
        """Create / refresh the CUDA engine from the current parameters and off-hooked tables."""
        _require_cuda("Hyper_SAGNN")
        emb = self.encode1.static_nn
        if len(emb.on_hook_set):
            emb.off_hook()
        if self._engine is not None:
            self._engine.close()
        dev = torch.cuda.current_device() if device_index is None else device_index
        eng = _engine.ScorerEngine(dev)
        num = [int(emb.num_list[i + 1] - emb.num_list[i]) for i in range(len(emb.wstack))]
        tables = [e.embedding for e in emb.off_hook_embedding]
        eng.load_model(self.state_dict(), num, tables=tables, cell_feats=self.cell_feats)
        if csr is not None:
            eng.set_contacts(csr)
            eng.set_neighbors(nbr, nbr_w)
        self._engine = eng
        self._prepared = None           # a new engine has not been prepared for any (activation, local_transfer_range) yet
        return eng

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engine"] = None            # the CUDA contexts are not part of the pickle (torch.save(model))
        for k in ("_teng", "_teng_key", "_tmap", "_tflat", "_tversion", "_tseed", "_prepared"):
            st.pop(k, None)
        return st

    # -- training forward (three heads) under torch autograd ------------------------------------------------------------
    def _train_engine(self, stage: int):
        """TrainEngine bound to this module's parameters (created lazily, rebuilt when the hook state changes)."""
        from ..train_engine import TrainEngine
        _require_cuda("Hyper_SAGNN.forward")
        emb = self.encode1.static_nn
        cell_on = 0 in emb.on_hook_set
        cell_table = None if cell_on else emb.off_hook_embedding[0].embedding
        key = (stage, cell_on, None if cell_table is None else cell_table.data_ptr(), torch.cuda.current_device())
        if getattr(self, "_teng", None) is not None and self._teng_key == key:
            return self._teng
        if getattr(self, "_teng", None) is not None:
            self._teng.close()
        num = [int(emb.num_list[i + 1] - emb.num_list[i]) for i in range(len(emb.wstack))]
        feats = [e.embedding.detach().cpu().numpy().astype(np.float32) for e in emb.embeddings]
        eng = TrainEngine(self.state_dict(), num, feats, self.cell_feats, device=torch.cuda.current_device(), cell_table=cell_table,
                          cell_on_hook=(cell_on and stage == 1))
        named = dict(self.named_parameters(remove_duplicate=False))
        self._tmap = [(name, named[name], o, int(np.prod(shape))) for name, (o, shape) in eng.layout.items() if name in named]
        self._tflat = torch.zeros(eng.total, dtype=torch.float32).pin_memory()
        self._tversion = None
        self._teng, self._teng_key, self._tseed = eng, key, 0
        return eng

    def _push_params(self, eng):
        version = tuple(p._version for _, p, _, _ in self._tmap)
        if version == self._tversion:
            return
        for _, prm, o, n in self._tmap:
            self._tflat[o:o + n].copy_(prm.detach().reshape(-1))
        eng.params.copy_(self._tflat, non_blocking=False)
        self._tversion = version

    def _pull_grads(self, eng, chrom, stage, zinb):
        from ..train_engine import grad_presence
        dev_flat = eng.grads.detach()
        host_flat = None                 # one device-to-host copy, and only when some parameter lives on the host
        names = [t[0] for t in self._tmap]
        for (name, prm, o, n), on in zip(self._tmap, grad_presence(names, chrom, stage, zinb, eng.cell_on_hook)):
            if not on or not prm.requires_grad:
                continue
            if prm.device == dev_flat.device:
                g = dev_flat[o:o + n].view(prm.shape)
            elif prm.device.type == "cpu":
                if host_flat is None:
                    host_flat = dev_flat.cpu()
                g = host_flat[o:o + n].view(prm.shape)
            else:
                g = dev_flat[o:o + n].view(prm.shape).to(prm.device)
            prm.grad = g.clone() if prm.grad is None else prm.grad + g

    def forward(self, x, x_chrom, mask=None, chroms_in_batch=None):
        """Modules.py:726-783: (mean, var, proba) heads [B,1] of a minibatch of ONE chromosome, differentiable: the forward and
        backward passes run in the CUDA training kernels (hsg_score_fwd / hsg_score_bwd); loss.backward() leaves the gradients
        in the .grad of this module's parameters, so the reference's train_epoch body works unchanged.  `x_chrom` is either the
        chromosome array or the (chromosome array, to_neighs) pair that forward_batch_hyperedge passes (Higashi_wrapper.py:867)."""
        to_neighs = None
        if isinstance(x_chrom, (tuple, list)) and len(x_chrom) == 2 and not np.isscalar(x_chrom[0]):
            x_chrom, to_neighs = x_chrom
        chroms = np.asarray(x_chrom.detach().cpu() if hasattr(x_chrom, "detach") else x_chrom).reshape(-1)
        if len(chroms) == 0 or np.any(chroms != chroms[0]):
            raise ValueError("one chromosome per minibatch (DataGenerator.next_iter with k=1, Modules.py:1205)")
        stage = 2 if isinstance(self.encode1.dynamic_nn, GraphSageEncoder_with_weights) else 1
        if stage == 2:
            if isinstance(to_neighs, list) and len(to_neighs) == 1:
                to_neighs = to_neighs[0]
            if to_neighs is None:
                raise ValueError("the GraphSAGE encoder needs the to_neighs_to_mask tuple: model(x, (chrom, to_neighs))")
        anchor = torch.zeros((), dtype=torch.float32, requires_grad=torch.is_grad_enabled())
        return _HyperedgeScoreFn.apply(anchor, self, x, int(chroms[0]), to_neighs, stage)

    def predict(self, input, input_chrom, verbose=False, batch_size=96, activation=None, extra_info=None):
        """Modules.py:786-827: scores of explicit triplets [N,3] (cell+1, bin_i, bin_j) for the cell fixed by
        `fix_cell2`; returns float32 [N,1] on the host.  `batch_size` is accepted for compatibility (one launch)."""
        _require_cuda("Hyper_SAGNN.predict")
        eng = self._engine
        if eng is None:
            raise RuntimeError("call build_engine(csr, ...) (or impute_process) before predict")
        dyn = self.encode1.dynamic_nn
        if getattr(dyn, "fixed_cell", None) is None:
            raise RuntimeError("call encode1.dynamic_nn.fix_cell2(cell, ...) first")
        act = {None: _engine.ACT_NONE, torch.sigmoid: _engine.ACT_SIGMOID,
               torch.nn.functional.softplus: _engine.ACT_SOFTPLUS}.get(activation, None)
        if act is None:
            raise ValueError("activation must be None, torch.sigmoid or F.softplus")
        key = (act, int(getattr(dyn, "local_transfer_range", 0)))
        if getattr(self, "_prepared", None) != key:
            eng.prepare(_engine.PREC_FP32, key[1], act)
            self._prepared = key
        x = torch.as_tensor(np.asarray(input), dtype=torch.int64).cuda().contiguous()
        out = torch.empty(x.shape[0], dtype=torch.float32, device="cuda")
        eng.fix_cell(dyn.fixed_cell)
        eng.score_triplets(x, out)
        res = out.cpu().numpy().reshape(-1, 1)
        if extra_info is not None:
            xi = np.asarray(input)
            res = res * np.asarray(extra_info)[xi[:, 2] - xi[:, 1]].reshape(-1, 1)
        return res


# ----------------------------------------------------------------------------- positive sampler (host)
class DataGenerator:
    """Chromosome-balanced positive sampler (Modules.py:1165-1234): one random chromosome per batch (k=1),
    pointer + reshuffle, `filter_edges(min_bin, max_bin)`."""

    def __init__(self, edges, edge_chrom, edge_weight, batch_size, flag=False, num_list=None, k=1):
        self.batch_size = int(batch_size)
        self.flag, self.k = flag, k
        self.num_list = list(num_list)
        self.edges, self.edge_weight, self.edge_chrom = list(edges), list(edge_weight), list(edge_chrom)
        self.chrom_list = np.arange(len(self.num_list) - 1)
        sizes = []
        for i in range(len(self.num_list) - 1):
            sizes.append(len(self.edges[i]))
            if len(self.edges[i]) == 0:
                continue
            while len(self.edges[i]) <= self.batch_size:
                self.edges[i] = np.concatenate([self.edges[i], self.edges[i]])
                self.edge_weight[i] = np.concatenate([self.edge_weight[i], self.edge_weight[i]])
                self.edge_chrom[i] = np.concatenate([self.edge_chrom[i], self.edge_chrom[i]])
        self.pointer = np.zeros(int(np.max(self.chrom_list) + 1)).astype("int")
        self.size_list = np.asarray(sizes, dtype=np.float64) / max(1, np.sum(sizes))

    def filter_edges(self, min_bin=0, max_bin=-1):
        for i in range(len(self.edges)):
            dist = self.edges[i][:, 2] - self.edges[i][:, 1]
            m = (dist > min_bin) & (dist < max_bin)
            self.edges[i], self.edge_weight[i], self.edge_chrom[i] = self.edges[i][m], self.edge_weight[i][m], self.edge_chrom[i][m]

    def next_iter(self):
        chroms = np.random.choice(self.chrom_list, size=self.k, replace=True)
        bs = int(self.batch_size / self.k)
        e_list, c_list, w_list = [], [], []
        for chrom in chroms:
            if len(self.edges[chrom]) == 0:
                continue
            self.pointer[chrom] += bs
            if self.pointer[chrom] > len(self.edges[chrom]):
                index = np.random.permutation(len(self.edges[chrom]))
                self.edges[chrom], self.edge_weight[chrom], self.edge_chrom[chrom] = \
                    self.edges[chrom][index], self.edge_weight[chrom][index], self.edge_chrom[chrom][index]
                self.pointer[chrom] = bs
            sl = slice(self.pointer[chrom] - bs, min(self.pointer[chrom], len(self.edges[chrom])))
            e_list.append(self.edges[chrom][sl]); c_list.append(self.edge_chrom[chrom][sl]); w_list.append(self.edge_weight[chrom][sl])
        return np.concatenate(e_list), np.concatenate(c_list), np.concatenate(w_list), chroms
