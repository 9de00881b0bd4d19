"""Literal CPU restatement of the Hyper-SAGNN scorer (torch-CPU fp32, eval mode).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function follows the reference
op for op (same batching, same per-triplet recomputation) so that it is both the
parity oracle and an honest CPU baseline:

    Hyper_SAGNN.forward          higashi/Higashi_backend/Modules.py:726-783
    Hyper_SAGNN.predict          Modules.py:786-827
    EncoderLayer.forward         Modules.py:1136-1161
    MultipleEmbedding.forward    Modules.py:570-602   (+ off_hook tables :605-631)
    TiedAutoEncoder.encoder      Modules.py:235-253
    GraphSageEncoder_with_weights.forward_on_hook / forward_off_hook / fix_cell2
                                 Modules.py:1485-1553 / 1472-1483 / 1432-1469
    MeanAggregator_with_weights.forward / forward_GCN   Modules.py:1337-1371
    MultiHeadAttention.forward   Modules.py:1029-1095
    ScaledDotProductAttention    Modules.py:936-972
    PositionwiseFeedForward      Modules.py:865-889
    FeedForward                  Modules.py:914-926
    swish                        higashi/Higashi_backend/Functions.py:14

`sd` is a dict name -> torch.Tensor with the REFERENCE's state_dict names.
"""
import math
import warnings

import numpy as np
import torch
import torch.nn.functional as F

warnings.filterwarnings("ignore", message="Sparse invariant checks")
N_HEAD, D_K, D_V = 8, 16, 16          # Higashi_wrapper.py:834-845


def swish(x):
    return x * torch.sigmoid(x)


def as_state(arrays, prefix="sd/"):
    """npz-style {prefix+name: ndarray} -> {name: tensor}."""
    return {k[len(prefix):]: torch.from_numpy(np.asarray(v)) for k, v in arrays.items() if k.startswith(prefix)}


# ----------------------------------------------------------------- small modules
def _ln(x, sd, name):
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def feed_forward(x, sd, name):
    """FeedForward: Linear stack, swish between layers (Modules.py:914-926)."""
    n = 0
    while "%s.w_stack.%d.weight" % (name, n) in sd:
        n += 1
    out = x
    for i in range(n):
        out = F.linear(out, sd["%s.w_stack.%d.weight" % (name, i)], sd.get("%s.w_stack.%d.bias" % (name, i)))
        if i < n - 1:
            out = swish(out)
    return out


def pff(x, sd, name, layer_norm=False, residual=False):
    """PositionwiseFeedForward: [LN] -> 1x1 conv stack with swish -> [+x] (Modules.py:865-889)."""
    n = 0
    while "%s.w_stack.%d.weight" % (name, n) in sd:
        n += 1
    out = _ln(x, sd, name + ".layer_norm") if layer_norm else x
    for i in range(n):
        w = sd["%s.w_stack.%d.weight" % (name, i)][:, :, 0]          # Conv1d k=1 == Linear
        out = F.linear(out, w, sd["%s.w_stack.%d.bias" % (name, i)])
        if i < n - 1:
            out = swish(out)
    if residual and out.shape[-1] == x.shape[-1]:
        out = out + x
    return out


def encoder_branch(feat, sd, prefix, i):
    """wstack[i] applied to raw node features: cells (i==0) linear, bins swish (Modules.py:550-558, 235-253)."""
    w = sd["%s.wstack.%d.weight_list.0" % (prefix, i)]
    b = sd["%s.wstack.%d.bias_list.0" % (prefix, i)]
    out = F.linear(feat, w, b)
    return swish(out) if i > 0 else out


def embed_tables(sd, feats, prefix="encode1.static_nn"):
    """off_hook(): precomputed [n_nodes_of_type, d] tables for every branch (Modules.py:605-631)."""
    return [encoder_branch(torch.as_tensor(f), sd, prefix, i) for i, f in enumerate(feats)]


def multiple_embedding(ids, num_list, sd, feats, tables=None, prefix="encode1.static_nn"):
    """MultipleEmbedding.forward for a flat id vector (Modules.py:570-602).

    `tables[i]` (if not None) replaces branch i by its off-hooked table."""
    ids = ids.reshape(-1)
    d = sd["%s.wstack.0.weight_list.0" % prefix].shape[0]
    final = torch.zeros((len(ids), d), dtype=torch.float32)
    bounds = [0] + [int(v) for v in num_list]
    for i in range(len(bounds) - 1):
        mask = (ids > bounds[i]) & (ids <= bounds[i + 1])
        if not bool(mask.any()):
            continue
        local = ids[mask] - bounds[i] - 1
        if tables is not None and tables[i] is not None:
            final[mask] = tables[i][local]
        else:
            final[mask] = encoder_branch(torch.as_tensor(feats[i])[local], sd, prefix, i)
    return final


# ----------------------------------------------------------------- attention block
def masked_attention(q, k, v):
    """(n*b) x 3 x 16 tensors; diagonal-masked softmax exactly as Modules.py:936-972."""
    attn = torch.bmm(q, k.transpose(1, 2)) / math.sqrt(D_K)
    L = q.shape[1]
    mask = (torch.ones(L, L) - torch.eye(L))[None]
    res = F.softmax(attn * mask, dim=-1)
    res = res * mask
    res = res / (res.sum(dim=-1, keepdim=True) + 1e-13)
    return torch.bmm(res, v)


def multi_head_attention(dynamic, static, sd, name="encode1.mul_head_attn"):
    """MultiHeadAttention.forward(q=dynamic, k=dynamic, v=static), eval mode (Modules.py:1029-1095)."""
    b, L, _ = dynamic.shape
    q = _ln(dynamic, sd, name + ".layer_norm1")
    k = _ln(dynamic, sd, name + ".layer_norm2")
    v = _ln(static, sd, name + ".layer_norm3")
    q = F.linear(q, sd[name + ".w_qs.weight"]).view(b, L, N_HEAD, D_K).permute(2, 0, 1, 3).reshape(-1, L, D_K)
    k = F.linear(k, sd[name + ".w_ks.weight"]).view(b, L, N_HEAD, D_K).permute(2, 0, 1, 3).reshape(-1, L, D_K)
    v = F.linear(v, sd[name + ".w_vs.weight"]).view(b, L, N_HEAD, D_V).permute(2, 0, 1, 3).reshape(-1, L, D_V)
    dyn = masked_attention(q, k, v)
    dyn = dyn.view(N_HEAD, b, L, D_V).permute(1, 2, 0, 3).reshape(b, L, -1)
    sta = v.view(N_HEAD, b, L, D_V).permute(1, 2, 0, 3).reshape(b, L, -1)
    dyn = feed_forward(dyn, sd, name + ".fc1")
    sta = feed_forward(sta, sd, name + ".fc2")
    return dyn, sta


def distance_side_channel(x, sd, cell_feats, only_model):
    """extra_proba{,2,3} on [attr[b_i] | attr[b_j] | cell_feats[cell]] (Modules.py:729-740)."""
    attr = sd["attribute_dict_embedding.weight"]
    cf = torch.as_tensor(cell_feats)
    c = torch.zeros((len(x), cf.shape[-1])) if only_model else cf[x[:, 0]]
    dist = torch.cat([attr[x[:, 1]], attr[x[:, 2]], c], dim=-1)
    return (feed_forward(dist, sd, "extra_proba"), feed_forward(dist, sd, "extra_proba2"),
            feed_forward(dist, sd, "extra_proba3"))


def heads(dynamic, static, x, sd, cell_feats, only_model):
    """Tail of Hyper_SAGNN.forward (Modules.py:752-783). Returns (mean, var, proba) each [B,1]."""
    dp, dp2, dp3 = distance_side_channel(x, sd, cell_feats, only_model)
    dynamic = _ln(dynamic, sd, "layer_norm1")
    static = _ln(static, sd, "layer_norm2")
    output = (dynamic - static) ** 2
    proba = pff(static, sd, "pff_classifier_proba").mean(dim=-2) + dp
    mean = pff(output, sd, "pff_classifier").mean(dim=-2) + dp2
    var = pff(static, sd, "pff_classifier_var").mean(dim=-2) + dp3
    return mean, var, proba


def encode_and_score(dynamic, static, x, sd, cell_feats, only_model):
    """EncoderLayer tail (MHA + pff_n1; pff_n2 is never called, Modules.py:1157-1161) + heads."""
    dyn, sta = multi_head_attention(dynamic, static, sd)
    dyn = pff(dyn, sd, "encode1.pff_n1", layer_norm=True, residual=True)
    return heads(dyn, sta, x, sd, cell_feats, only_model)


# ----------------------------------------------------------------- the three model states
def forward_stage1(x, num_list, sd, feats, cell_feats, tables=None, only_model=False):
    """dynamic_nn = static_nn = MultipleEmbedding (Modules.py:1152-1155)."""
    x = torch.as_tensor(x).long()
    B = len(x)
    static = multiple_embedding(x.reshape(-1), num_list, sd, feats, tables).view(B, 3, -1)
    return encode_and_score(static, static, x, sd, cell_feats, only_model)


def forward_stage2(x, to_neighs, num_list, sd, feats, cell_feats, tables, only_model=False):
    """GraphSAGE encoder on-hook (Modules.py:1485-1553, 1337-1362).

    to_neighs = (indices int64 [2,nnz], values f32 [nnz], unique_nodes int64 [U]); rows 2r, 2r+1
    are the two bins of hyperedge r.  tables[0] is the frozen cell table (off_hook([0]))."""
    x = torch.as_tensor(x).long()
    B = len(x)
    ind, val, uniq = [torch.as_tensor(t) for t in to_neighs]
    bins = x[:, 1:].reshape(-1)
    cell = multiple_embedding(x[:, 0], num_list, sd, feats, tables)
    emb_u = multiple_embedding(uniq.long(), num_list, sd, feats, tables)
    mask = torch.sparse_coo_tensor(ind.long(), val.float(), (len(bins), len(uniq)))
    neigh = torch.sparse.mm(mask, emb_u) if len(uniq) else torch.zeros((len(bins), emb_u.shape[-1]))
    self_f = multiple_embedding(bins, num_list, sd, feats, tables)
    comb = swish(F.linear(torch.cat([neigh, self_f], dim=-1), sd["encode1.dynamic_nn.nn.weight"],
                          sd["encode1.dynamic_nn.nn.bias"]))
    dynamic = torch.cat([cell[:, None, :], comb.view(B, 2, -1)], dim=1)
    static = torch.cat([cell[:, None, :], self_f.view(B, 2, -1)], dim=1)
    return encode_and_score(dynamic, static, x, sd, cell_feats, only_model)


def fix_cell(adj_list, tables, num_list):
    """fix_cell2 (Modules.py:1432-1469): bin_feats[global bin] = [A_hat . E_c | E_c] per chromosome.

    adj_list[c] = normalised adjacency of the cell for chromosome c (scipy sparse or dense)."""
    d = tables[0].shape[-1]
    bin_feats = torch.zeros((int(num_list[-1]) + 1, 2 * d), dtype=torch.float32)
    for c, adj in enumerate(adj_list):
        E = tables[c + 1]
        coo = adj.tocoo()
        A = torch.sparse_coo_tensor(np.stack([coo.row, coo.col]).astype(np.int64), torch.from_numpy(coo.data.astype(np.float32)),
                                    coo.shape)
        neigh = torch.sparse.mm(A, E)
        lo = int(num_list[c]) + 1
        bin_feats[lo:lo + E.shape[0]] = torch.cat([neigh, E], dim=-1)
    return bin_feats


def forward_offhook(x, bin_feats, num_list, sd, tables, cell_feats, only_model=True):
    """forward_off_hook (Modules.py:1472-1483) + scorer, per triplet as the reference does."""
    x = torch.as_tensor(x).long()
    B = len(x)
    bins = x[:, 1:].reshape(-1)
    cell = tables[0][x[:, 0] - 1]
    dyn_b = swish(F.linear(bin_feats[bins], sd["encode1.dynamic_nn.nn.weight"], sd["encode1.dynamic_nn.nn.bias"]))
    sta_b = multiple_embedding(bins, num_list, sd, None, tables)
    dynamic = torch.cat([cell[:, None, :], dyn_b.view(B, 2, -1)], dim=1)
    static = torch.cat([cell[:, None, :], sta_b.view(B, 2, -1)], dim=1)
    return encode_and_score(dynamic, static, x, sd, cell_feats, only_model)


def predict(x, bin_feats, num_list, sd, tables, cell_feats, activation="sigmoid", batch_size=100000):
    """Hyper_SAGNN.predict batch loop (Modules.py:786-827): mean head -> activation."""
    act = torch.sigmoid if activation == "sigmoid" else F.softplus
    outs = []
    with torch.no_grad():
        for j in range(math.ceil(len(x) / batch_size)):
            xb = x[j * batch_size:(j + 1) * batch_size]
            m, _, _ = forward_offhook(xb, bin_feats, num_list, sd, tables, cell_feats, only_model=True)
            outs.append(act(m))
    return torch.cat(outs, dim=0).numpy() if outs else np.zeros((0, 1), np.float32)
