"""Mint the golden vectors by running the UNMODIFIED reference.

Run by hand in the build container (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference through `_ref_shims`, builds small seeded synthetic
inputs, runs the reference functions named below and stores inputs + outputs
as `tests/golden/*.npz`.  The fixtures are committed; this script never runs on
the GPU box.

Fixture              reference entry point
-------------------  ----------------------------------------------------------
binpair.npz          Impute.generate_binpair                     (Impute.py:49)
knn.npz              Higashi.get_cell_neighbor        (Higashi_wrapper.py:1303)
neighs.npz           one_thread_generate_neg / to_neighs_to_mask (wrapper:236,171)
scorer_d{64,128}.npz Hyper_SAGNN.forward (stage 1, stage 2 on-hook) and
                     .predict after fix_cell2 (off-hook)      (Modules.py:726,786)
impute_*.npz         Impute.impute_process end to end            (Impute.py:123)
grads_d64.npz        autograd gradients of one eval-mode batch
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _ref_shims  # noqa: E402

RM, RI, RW = _ref_shims.import_reference()
import torch  # noqa: E402

from higashi_b200 import synthetic as S  # noqa: E402

torch.set_num_threads(4)


# ------------------------------------------------------------------ helpers
def tiny_dataset(seed=11, n_cells=24, bins=(40, 33), density=0.08, F=12):
    ds = S.make_dataset("T", n_cells=n_cells, bins=list(bins), seed=seed, density=density,
                        with_features=True)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    ds.cell_embed_feats = rng.standard_normal((n_cells, F)).astype(np.float32)
    return ds


def build_reference_model(ds, d, seed=0):
    """Appendix A of SURVEY.md: stage-1 model with randomised affine parameters."""
    torch.manual_seed(seed)
    num_list = ds.num_list
    emb = [ds.cell_embed_feats] + list(ds.bin_feats)
    node_embedding_init = RM.MultipleEmbedding(emb, d, False, num_list, emb)
    model = RM.Hyper_SAGNN(n_head=8, d_model=d, d_k=16, d_v=16, diag_mask=True, bottle_neck=d,
                           attribute_dict=ds.attribute_dict, cell_feats=ds.cell_feats,
                           encoder_dynamic_nn=node_embedding_init,
                           encoder_static_nn=node_embedding_init, chrom_num=len(ds.bins))
    g = torch.Generator().manual_seed(seed + 123)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if not p.requires_grad:
                continue
            if "layer_norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif "layer_norm" in name and name.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias") or "bias_list" in name:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return model, node_embedding_init


def to_stage2(model, node_embedding_init, ds, d, seed=5):
    """wrapper:1384,1415-1424: freeze the cell branch, install the GraphSAGE encoder."""
    torch.manual_seed(seed)
    node_embedding_init.off_hook([0])
    n_nodes = int(ds.num_list[-1])
    start_end = np.zeros((n_nodes + 1, 2), dtype="int")
    for j in range(len(ds.bins)):
        start_end[ds.num_list[j] + 1: ds.num_list[j + 1] + 1] = (ds.num_list[j], ds.num_list[j + 1])
    start_end[1: ds.n_cells + 1] = (0, ds.n_cells)
    enc = RM.GraphSageEncoder_with_weights(features=node_embedding_init, linear_features=node_embedding_init,
                                           feature_dim=d, embed_dim=d, num_sample=8, gcn=False,
                                           num_list=ds.num_list, transfer_range=0, start_end_dict=start_end,
                                           pass_pseudo_id=False, remove=True, pass_remove=False)
    with torch.no_grad():
        enc.nn.bias.add_(0.1 * torch.randn(enc.nn.bias.shape))
    model.encode1.dynamic_nn = enc
    return enc, start_end


def state_arrays(model):
    out = {}
    for k, v in model.state_dict().items():
        if k.startswith("encode1.dynamic_nn.features.") or k.startswith("encode1.dynamic_nn.linear_features.") \
                or k.startswith("encode1.dynamic_nn.aggregator.") or k.startswith("encode1.dynamic_nn.linear_aggregator."):
            continue        # aliases of encode1.static_nn.*
        out["sd/" + k] = v.detach().cpu().numpy()
    return out


def dataset_arrays(ds):
    out = {"ds/num": ds.num, "ds/row_ptr": ds.csr.row_ptr, "ds/col": ds.csr.col, "ds/val": ds.csr.val,
           "ds/attribute_dict": ds.attribute_dict, "ds/cell_feats": ds.cell_feats,
           "ds/cell_embed_feats": ds.cell_embed_feats}
    for j, f in enumerate(ds.bin_feats):
        out["ds/bin_feats_%d" % j] = f
    return out


def set_wrapper_globals(ds, sparse_obj, start_end, neg_num=3, weighted=False, precompute=True,
                        cell_neighbor_list=None, weight_dict=None, steps=2, pair_ratio=0.5):
    RW.neg_num = neg_num
    RW.max_bin = int(max(ds.bins))
    RW.mode = "classification"
    RW.graphsagemode = True
    RW.precompute_weighted_nbr = precompute
    RW.sparse_chrom_list_GCN = sparse_obj
    RW.sparse_chrom_list_dict = sparse_obj
    RW.num_list = ds.num_list
    RW.start_end_dict = start_end
    RW.mem_efficient_flag = True
    RW.steps = steps
    RW.pair_ratio = pair_ratio
    RW.weighted_adj = weighted
    if cell_neighbor_list is not None:
        RW.cell_neighbor_list = cell_neighbor_list
        RW.weight_dict = weight_dict


def positives(ds, chrom, n, rng):
    """n positive (cell, bin_i, bin_j) hyperedges of one chromosome, sorted ids (Process.py:795-797)."""
    out = []
    while len(out) < n:
        cell = int(rng.integers(0, ds.n_cells))
        m = ds.csr.matrix(chrom, cell).tocoo()
        ok = (m.col - m.row) >= 2
        if ok.sum() == 0:
            continue
        t = int(rng.integers(0, ok.sum()))
        r, c = m.row[ok][t], m.col[ok][t]
        out.append([cell + 1, ds.num_list[chrom] + r + 1, ds.num_list[chrom] + c + 1])
    return np.array(out, dtype="int")


def knn_reference(cell_embeddings, k):
    ns = type("NS", (), {})()
    ns.cell_embeddings = cell_embeddings
    ns.pre_cell_embed = False
    ns.neighbor_num = k + 1
    ns.num = np.array([len(cell_embeddings)])
    RW.tqdm = lambda x, **kw: x
    nl, wl = RW.Higashi.get_cell_neighbor(ns, 0)
    return nl, wl


# ------------------------------------------------------------------ (i) binpair
def golden_binpair():
    out = {}
    cases = {
        "full_250": dict(start=0, end=250, min_bin_=1, max_bin_=int(1e5), skip=[]),
        "min0_60": dict(start=0, end=60, min_bin_=0, max_bin_=int(1e5), skip=[]),
        "offset_band": dict(start=620, end=620 + 180, min_bin_=2, max_bin_=25, skip=[]),
        "skip_band": dict(start=100, end=100 + 300, min_bin_=1, max_bin_=40, skip=list(range(220, 236))),
        "skip_edges": dict(start=0, end=90, min_bin_=1, max_bin_=int(1e5), skip=[0, 1, 2, 50, 51, 88, 89]),
        "banded_4864": dict(start=4000, end=4000 + 4864, min_bin_=1, max_bin_=200,
                            skip=list(range(4000 + 1830, 4000 + 1920))),
        "empty": dict(start=10, end=11, min_bin_=1, max_bin_=50, skip=[]),
    }
    for name, c in cases.items():
        res = RI.generate_binpair(c["start"], c["end"], c["min_bin_"], c["max_bin_"], set(c["skip"]))
        res = np.asarray(res)
        if res.ndim != 2:
            res = np.zeros((0, 2), dtype=np.int64)
        out[name + "/args"] = np.array([c["start"], c["end"], c["min_bin_"], c["max_bin_"]], dtype=np.int64)
        out[name + "/skip"] = np.array(c["skip"], dtype=np.int64)
        res = res.astype(np.int32)
        if len(res) > 100000:       # keep the fixture small: digest + head/tail rows
            import hashlib
            out[name + "/count"] = np.array(len(res))
            out[name + "/sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(res).tobytes()).hexdigest())
            out[name + "/head"] = res[:256]
            out[name + "/tail"] = res[-256:]
        else:
            out[name + "/pairs"] = res
    np.savez_compressed(os.path.join(HERE, "binpair.npz"), **out)
    print("binpair.npz", {k: v.shape for k, v in out.items() if k.endswith("pairs") or k.endswith("count")})


# ------------------------------------------------------------------ (ii) kNN
def golden_knn():
    out = {}
    for tag, (n, d, k, seed) in {"a": (620, 64, 4, 1), "b": (200, 32, 5, 2), "c": (31, 8, 4, 3)}.items():
        rng = np.random.Generator(np.random.PCG64(seed))
        v = rng.standard_normal((n, d)).astype(np.float32)
        v[:, :3] += rng.integers(0, 4, size=(n, 1)) * 2.0          # some cluster structure
        nl, wl = knn_reference(v, k)
        out[tag + "/embed"] = v
        out[tag + "/k"] = np.array(k)
        out[tag + "/nbr"] = np.stack([np.asarray(x) for x in nl[1:]]).astype(np.int64)
        out[tag + "/w"] = np.stack([np.asarray(x) for x in wl[1:]])
    np.savez_compressed(os.path.join(HERE, "knn.npz"), **out)
    print("knn.npz", out["a/nbr"].shape, out["a/w"].dtype)


# ------------------------------------------------------------------ (iii) neighbour lists
def golden_neighs():
    ds = tiny_dataset()
    sparse_obj = ds.csr.to_object_array()
    model, nei = build_reference_model(ds, 64)
    enc, start_end = to_stage2(model, nei, ds, 64)
    out = dataset_arrays(ds)
    rng = np.random.Generator(np.random.PCG64(77))
    emb = rng.standard_normal((ds.n_cells, 8)).astype(np.float32)
    nl, wl = knn_reference(emb, 3)
    weight_dict = {}
    for i in range(len(nl)):
        for c, w in zip(nl[i], wl[i]):
            weight_dict[(c, i)] = w
    out["knn/nbr"] = np.stack([np.asarray(x) for x in nl[1:]]).astype(np.int64)
    out["knn/w"] = np.stack([np.asarray(x) for x in wl[1:]]).astype(np.float32)

    old_perm, old_rng = np.random.permutation, np.random.default_rng
    for tag, weighted, training in [("plain_eval", False, False), ("plain_train", False, True),
                                    ("knn_eval", True, False), ("knn_train", True, True)]:
        chrom = 1 if "knn" in tag else 0
        edges = positives(ds, chrom, 12, rng)
        set_wrapper_globals(ds, sparse_obj, start_end, neg_num=2, weighted=weighted, precompute=False,
                            cell_neighbor_list=nl, weight_dict=weight_dict)
        np.random.permutation = lambda n: np.arange(n)           # keep the batch order
        RW.np.random.default_rng = lambda: np.random.Generator(np.random.PCG64(99))
        np.random.seed(5)
        try:
            x, y, w, x_chrom, to_neighs, _ = RW.one_thread_generate_neg(
                edges, np.full(len(edges), chrom, dtype="int8"), np.ones(len(edges), dtype="float32"),
                1, training, np.array([chrom]))
        finally:
            np.random.permutation = old_perm
            RW.np.random.default_rng = old_rng
        # the removal draws the reference used (same seeded generator, same call order:
        # generate_negative_cpu makes its own generator first; the one used for removal
        # is a fresh PCG64(99) whose first call is random(2*len(x)))
        rem = np.random.Generator(np.random.PCG64(99)).random(2 * len(x))
        ind, v, uniq = to_neighs[0]
        out[tag + "/x"] = x.astype(np.int64)
        out[tag + "/x_chrom"] = x_chrom.astype(np.int64)
        out[tag + "/y"] = y
        out[tag + "/remove_draw"] = rem
        out[tag + "/training"] = np.array(training)
        out[tag + "/indices"] = ind.numpy().astype(np.int64)
        out[tag + "/values"] = v.numpy()
        out[tag + "/unique"] = uniq.numpy().astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "neighs.npz"), **out)
    print("neighs.npz", {k: v.shape for k, v in out.items() if "indices" in k})


# ------------------------------------------------------------------ (iv) scorer outputs
def golden_scorer(d):
    ds = tiny_dataset()
    sparse_obj = ds.csr.to_object_array()
    rng = np.random.Generator(np.random.PCG64(1234 + d))
    out = dataset_arrays(ds)

    # ---- stage 1: MultipleEmbedding on-hook for both branches (wrapper:1327)
    model, nei = build_reference_model(ds, d)
    model.eval()
    for chrom in (0, 1):
        x = np.concatenate([positives(ds, chrom, 20, rng)])
        x[10:, 1] = np.minimum(x[10:, 1] + 1, x[10:, 2] - 1)       # a few arbitrary (negative-like) triplets
        with torch.no_grad():
            m, v, p = model(torch.from_numpy(x), (np.full(len(x), chrom, dtype="int8"), []),
                            chroms_in_batch=np.array([chrom]) + 1)
        out["stage1/c%d/x" % chrom] = x.astype(np.int64)
        out["stage1/c%d/out" % chrom] = torch.cat([m, v, p], dim=1).numpy()
    out.update({"stage1/" + k: v for k, v in state_arrays(model).items()})

    # ---- stage 2: GraphSAGE encoder on-hook, eval mode (wrapper:1377)
    enc, start_end = to_stage2(model, nei, ds, d)
    model.eval()
    out["stage2/cell_table"] = nei.off_hook_embedding[0].embedding.detach().numpy()
    old_perm, old_rng = np.random.permutation, np.random.default_rng
    for chrom in (0, 1):
        edges = positives(ds, chrom, 16, rng)
        set_wrapper_globals(ds, sparse_obj, start_end, neg_num=2, weighted=False, precompute=True)
        np.random.permutation = lambda n: np.arange(n)
        RW.np.random.default_rng = lambda: np.random.Generator(np.random.PCG64(7 + chrom))
        try:
            x, y, w, x_chrom, to_neighs, _ = RW.one_thread_generate_neg(
                edges, np.full(len(edges), chrom, dtype="int8"), np.ones(len(edges), dtype="float32"),
                1, False, np.array([chrom]))
        finally:
            np.random.permutation = old_perm
            RW.np.random.default_rng = old_rng
        with torch.no_grad():
            m, v, p = model(torch.from_numpy(x), (x_chrom, to_neighs[0]), chroms_in_batch=np.array([chrom]) + 1)
        ind, val, uniq = to_neighs[0]
        out["stage2/c%d/x" % chrom] = x.astype(np.int64)
        out["stage2/c%d/indices" % chrom] = ind.numpy().astype(np.int64)
        out["stage2/c%d/values" % chrom] = val.numpy()
        out["stage2/c%d/unique" % chrom] = uniq.numpy().astype(np.int64)
        out["stage2/c%d/out" % chrom] = torch.cat([m, v, p], dim=1).numpy()
    out.update({"stage2/" + k: v for k, v in state_arrays(model).items()})

    # ---- imputation mode: off-hook tables + fix_cell2 + predict (Impute.py:163-272)
    model.eval()
    model.only_model = True
    nei.off_hook()
    with torch.no_grad():
        enc.start_fix()
    enc.forward = enc.forward_off_hook
    RI.sparse_chrom_list = sparse_obj
    chrom_list = ["chr%d" % (j + 1) for j in range(len(ds.bins))]
    bin_ids = [np.arange(ds.num_list[j], ds.num_list[j + 1]) + 1 for j in range(len(ds.bins))]
    big, big_chrom = [], []
    for j in range(len(ds.bins)):
        sb = RI.generate_binpair(ds.num_list[j], ds.num_list[j + 1], 1, int(1e5), set())
        big.append(np.concatenate([np.ones((len(sb), 1), dtype="int"), sb], axis=-1))
        big_chrom.append(np.full(len(sb), j, dtype="int"))
    big = np.concatenate(big)
    big_chrom = np.concatenate(big_chrom)
    for act_name, act in (("sigmoid", torch.sigmoid), ("softplus", torch.nn.functional.softplus)):
        res = []
        for cell in (1, 7, ds.n_cells):
            with torch.no_grad():
                _, ccl, _ = RI.prep_one(False, chrom_list, chrom_list, 0, cell)
                enc.fix_cell2(cell, bin_ids, ccl, 0, route_nn_list=[j + 1 for j in range(len(ds.bins))])
                big[:, 0] = cell
                pr = model.predict(big, big_chrom, verbose=False, batch_size=700, activation=act)
            res.append(pr.reshape(-1))
        out["impute/%s" % act_name] = np.stack(res)
    out["impute/cells"] = np.array([1, 7, ds.n_cells])
    out["impute/pairs"] = big[:, 1:].astype(np.int64)
    out["impute/cell_table"] = enc.cell_feats.detach().numpy()
    for j in range(len(ds.bins)):
        out["impute/bin_table_%d" % j] = nei.off_hook_embedding[j + 1].embedding.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "scorer_d%d.npz" % d), **out)
    print("scorer_d%d.npz" % d, out["impute/sigmoid"].shape, out["stage2/c0/out"][:2])


# ------------------------------------------------------------------ (v) impute_process end to end
def golden_impute(tag, k, ltr, with_skip=False, mode="classification", band=None):
    ds = tiny_dataset(seed=21 + k, n_cells=20, bins=(36, 30, 17))
    sparse_obj = ds.csr.to_object_array()
    d = 64
    model, nei = build_reference_model(ds, d, seed=3)
    enc, start_end = to_stage2(model, nei, ds, d)
    sd = state_arrays(model)
    tmp = tempfile.mkdtemp(prefix="hsg_golden_")
    chrom_list = ["chr%d" % (j + 1) for j in range(len(ds.bins))]
    res = 1000000
    cfg = dict(resolution=res, impute_list=chrom_list, chrom_list=chrom_list, temp_dir=tmp,
               minimum_distance=res, maximum_distance=-1, local_transfer_range=ltr,
               embedding_name="emb", impute_verbose=0)
    if band is not None:
        cfg["minimum_impute_distance"] = band[0] * res
        cfg["maximum_impute_distance"] = band[1] * res
    skip_tab = np.zeros((0, 3), dtype=np.int64)
    if with_skip:
        cyto = os.path.join(tmp, "cyto.txt")
        with open(cyto, "w") as f:
            f.write("chr1\t0\t9000000\tp1\tgneg\n")
            f.write("chr1\t12300000\t14100000\tp11\tacen\n")
            f.write("chr1\t14100000\t15800000\tq11\tacen\n")
            f.write("chr2\t5000000\t6500000\tp11\tacen\n")
            f.write("chr3\t16400000\t16900000\tq11\tacen\n")
        cfg["cytoband_path"] = cyto
    os.makedirs(os.path.join(tmp, "embed"))
    rng = np.random.Generator(np.random.PCG64(5))
    cell_emb = rng.standard_normal((ds.n_cells, 8)).astype(np.float32)
    np.save(os.path.join(tmp, "embed", "emb_0_origin.npy"), cell_emb)
    cfg_path = os.path.join(tmp, "config.json")
    json.dump(cfg, open(cfg_path, "w"))
    _ref_shims.H5_STORE[os.path.join(tmp, "node_feats.hdf5")] = {"num": ds.num}
    sparse_path = os.path.join(tmp, "sparse_nondiag_adj_nbr_1.npy")
    np.save(sparse_path, sparse_obj, allow_pickle=True)
    winfo = None
    out = dataset_arrays(ds)
    if k > 0:
        nl, wl = knn_reference(cell_emb, k)
        wd = {}
        for i in range(len(nl)):
            for c, w in zip(nl[i], wl[i]):
                wd[(c, i)] = w
        winfo = os.path.join(tmp, "weighted_info.npy")
        np.save(winfo, np.array([nl, wd], dtype=object), allow_pickle=True)
        out["knn/nbr"] = np.stack([np.asarray(x) for x in nl[1:]]).astype(np.int64)
        out["knn/w"] = np.stack([np.asarray(x) for x in wl[1:]]).astype(np.float32)
    name = "emb_nbr_%d_impute" % k
    cell_start, cell_end = 2, 11
    RI.impute_process(cfg_path, model, name, mode, cell_start, cell_end, sparse_path, winfo)
    for j, chrom in enumerate(chrom_list):
        f = _ref_shims.H5_STORE[os.path.join(tmp, "%s_%s.hdf5" % (chrom, name))]
        out["out/%s/coordinates" % chrom] = np.asarray(f["coordinates"]).astype(np.int64)
        out["out/%s/cells" % chrom] = np.stack([np.asarray(f["cell_%d" % c]) for c in range(cell_start, cell_end)])
        if with_skip:
            s, e = RI.skip_start_end(cfg, chrom)
            for a, b in zip(s, e):
                skip_tab = np.concatenate([skip_tab, [[j, a, b]]])
    out["cfg/k"] = np.array(k)
    out["cfg/ltr"] = np.array(ltr)
    out["cfg/mode"] = np.array(mode)
    out["cfg/cell_range"] = np.array([cell_start, cell_end])
    out["cfg/skip"] = skip_tab
    out["cfg/band"] = np.array(band if band is not None else [1, -1])
    out.update(sd)
    np.savez_compressed(os.path.join(HERE, "impute_%s.npz" % tag), **out)
    print("impute_%s.npz" % tag, {c: out["out/%s/cells" % c].shape for c in chrom_list})


# ------------------------------------------------------------------ (vi) gradients
def golden_grads():
    d = 64
    ds = tiny_dataset()
    sparse_obj = ds.csr.to_object_array()
    rng = np.random.Generator(np.random.PCG64(4321))
    model, nei = build_reference_model(ds, d)
    enc, start_end = to_stage2(model, nei, ds, d)
    model.eval()                                   # dropout off => deterministic gradients
    out = dataset_arrays(ds)
    chrom = 0
    edges = positives(ds, chrom, 16, rng)
    set_wrapper_globals(ds, sparse_obj, start_end, neg_num=2, weighted=False, precompute=True)
    old_perm, old_rng = np.random.permutation, np.random.default_rng
    np.random.permutation = lambda n: np.arange(n)
    RW.np.random.default_rng = lambda: np.random.Generator(np.random.PCG64(17))
    try:
        x, y, w, x_chrom, to_neighs, _ = RW.one_thread_generate_neg(
            edges, np.full(len(edges), chrom, dtype="int8"), rng.random(len(edges)).astype("float32") + 0.5,
            1, False, np.array([chrom]))
    finally:
        np.random.permutation = old_perm
        RW.np.random.default_rng = old_rng
    m, v, p = model(torch.from_numpy(x), (x_chrom, to_neighs[0]), chroms_in_batch=np.array([chrom]) + 1)
    yt, wt = torch.from_numpy(y).float(), torch.from_numpy(w).float()
    loss = torch.nn.functional.binary_cross_entropy_with_logits(m, yt, weight=wt)
    loss.backward()
    ind, val, uniq = to_neighs[0]
    out["x"] = x.astype(np.int64); out["y"] = y; out["w"] = w
    out["indices"] = ind.numpy().astype(np.int64); out["values"] = val.numpy(); out["unique"] = uniq.numpy().astype(np.int64)
    out["out"] = torch.cat([m, v, p], dim=1).detach().numpy()
    out["loss"] = np.array(loss.item())
    seen = set()
    for name, prm in model.named_parameters():
        if name.startswith("encode1.dynamic_nn.features.") or name.startswith("encode1.dynamic_nn.linear_features.") \
                or "aggregator" in name:
            continue
        if prm.grad is not None:
            out["grad/" + name] = prm.grad.detach().numpy()
        seen.add(name)
    # the shared MultipleEmbedding parameters are first registered under encode1.dynamic_nn.features.*;
    # store their gradients under the canonical encode1.static_nn.* names
    for name, prm in model.encode1.static_nn.named_parameters():
        if prm.grad is not None:
            out["grad/encode1.static_nn." + name] = prm.grad.detach().numpy()
    out.update(state_arrays(model))
    np.savez_compressed(os.path.join(HERE, "grads_d64.npz"), **out)
    print("grads_d64.npz loss", loss.item(), "n grads", sum(k.startswith("grad/") for k in out))


# ------------------------------------------------------------------ (vii) loss modes + stage-1 reconstruction
def golden_losses():
    """forward_batch_hyperedge (Higashi_wrapper.py:861-913) in zinb / rank / regression mode on one stage-2 batch, and the
    stage-1 batch with the cell auto-encoder reconstruction loss (use_recon=True); autograd gradients of each loss."""
    d = 64
    ds = tiny_dataset()
    sparse_obj = ds.csr.to_object_array()
    rng = np.random.Generator(np.random.PCG64(999))
    out = dataset_arrays(ds)

    def batch(chrom, seed, n=14):
        edges = positives(ds, chrom, n, rng)
        old_perm, old_rng = np.random.permutation, np.random.default_rng
        np.random.permutation = lambda k: np.arange(k)
        RW.np.random.default_rng = lambda: np.random.Generator(np.random.PCG64(seed))
        np.random.seed(seed)
        try:
            res = RW.one_thread_generate_neg(edges, np.full(len(edges), chrom, dtype="int8"),
                                             (rng.random(len(edges)) * 3 + 0.5).astype("float32"), 1, False, np.array([chrom]))
        finally:
            np.random.permutation = old_perm
            RW.np.random.default_rng = old_rng
        return res

    def grads_of(model, extra=()):
        g = {}
        for name, prm in list(model.named_parameters()):
            if name.startswith("encode1.dynamic_nn.features.") or name.startswith("encode1.dynamic_nn.linear_features.") or "aggregator" in name:
                continue
            if prm.grad is not None:
                g[name] = prm.grad.detach().numpy().copy()
        for name, prm in model.encode1.static_nn.named_parameters():
            if prm.grad is not None:
                g["encode1.static_nn." + name] = prm.grad.detach().numpy().copy()
        return g

    cell2weight = (rng.random(ds.n_cells) + 0.5).astype(np.float32)
    out["cell2weight"] = cell2weight
    # ---- stage 2, three loss modes
    model, nei = build_reference_model(ds, d)
    enc, start_end = to_stage2(model, nei, ds, d)
    model.eval()
    set_wrapper_globals(ds, sparse_obj, start_end, neg_num=2, weighted=False, precompute=True)
    x, y, w, x_chrom, to_neighs, _ = batch(1, 31)
    # mimic the mode-specific weights: zinb / regression use raw counts as targets (zeros for negatives)
    w = w.astype(np.float32)
    ind, val, uniq = to_neighs[0]
    out["s2/x"] = x.astype(np.int64); out["s2/y"] = y; out["s2/w"] = w
    out["s2/indices"] = ind.numpy().astype(np.int64); out["s2/values"] = val.numpy(); out["s2/unique"] = uniq.numpy().astype(np.int64)
    ns = type("NS", (), {})()
    ns.higashi_model = model; ns.use_recon = False; ns.node_embedding_init = nei
    ns.cell_feats1 = torch.from_numpy(cell2weight); ns.rank_thres = 1
    RW.neg_num = 2
    for mode in ("zinb", "rank", "regression", "classification"):
        ns.mode = mode
        model.zero_grad()
        pred, main_loss, mse_loss = RW.Higashi.forward_batch_hyperedge(
            ns, torch.from_numpy(x), torch.from_numpy(w).float(), x_chrom, to_neighs[0], torch.from_numpy(y).float(), np.array([1]))
        main_loss.backward()     # with use_recon=False train_epoch weights mse_loss by ratio1 = 0 (Higashi_wrapper.py:997-1002)
        out["s2/%s/loss" % mode] = np.array(float(main_loss))
        out["s2/%s/mse" % mode] = np.array(float(mse_loss.sum()))
        out["s2/%s/pred" % mode] = pred.detach().numpy()
        with torch.no_grad():
            m, v, p = model(torch.from_numpy(x), (x_chrom, to_neighs[0]), chroms_in_batch=np.array([1]) + 1)
        out["s2/%s/heads" % mode] = torch.cat([m, v, p], dim=1).numpy()
        for k, g in grads_of(model).items():
            out["s2/%s/grad/%s" % (mode, k)] = g
    out.update({"s2/" + k: v for k, v in state_arrays(model).items()})

    # ---- stage 1 with the reconstruction loss
    model1, nei1 = build_reference_model(ds, d, seed=8)
    model1.eval()
    x1 = positives(ds, 0, 18, rng)
    x1[9:, 2] = np.minimum(x1[9:, 2] + 1, ds.num_list[1])
    y1 = np.concatenate([np.ones((9, 1)), np.zeros((9, 1))]).astype(np.float32)
    w1 = (rng.random((18, 1)) + 0.5).astype(np.float32)
    ns1 = type("NS", (), {})()
    ns1.higashi_model = model1; ns1.use_recon = True; ns1.node_embedding_init = nei1; ns1.mode = "classification"
    ns1.cell_ids = torch.arange(ds.n_cells).long()
    model1.zero_grad()
    pred, main_loss, mse_loss = RW.Higashi.forward_batch_hyperedge(
        ns1, torch.from_numpy(x1), torch.from_numpy(w1), np.zeros(18, dtype="int8"), [], torch.from_numpy(y1), np.array([0]))
    main_loss.backward(retain_graph=True)
    gm = grads_of(model1)
    model1.zero_grad()
    mse_loss.backward()
    gr = grads_of(model1)
    out["s1/x"] = x1.astype(np.int64); out["s1/y"] = y1; out["s1/w"] = w1
    out["s1/loss"] = np.array(float(main_loss)); out["s1/mse"] = np.array(float(mse_loss)); out["s1/pred"] = pred.detach().numpy()
    for k, g in gm.items():
        out["s1/grad_main/" + k] = g
    for k, g in gr.items():
        out["s1/grad_recon/" + k] = g
    out.update({"s1/" + k: v for k, v in state_arrays(model1).items()})
    np.savez_compressed(os.path.join(HERE, "losses_d64.npz"), **out)
    print("losses_d64.npz", {m: float(out["s2/%s/loss" % m]) for m in ("zinb", "rank", "regression", "classification")},
          "stage1", float(out["s1/loss"]), float(out["s1/mse"]), "n arrays", len(out))


if __name__ == "__main__":
    golden_binpair()
    golden_knn()
    golden_neighs()
    golden_scorer(64)
    golden_scorer(128)
    golden_impute("k0_r0", k=0, ltr=0)
    golden_impute("k4_r0", k=4, ltr=0, with_skip=True)
    golden_impute("k4_r1", k=4, ltr=1, mode="zinb")
    golden_impute("k0_r1_band", k=0, ltr=1, band=(2, 12))
    golden_grads()
    golden_losses()
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print("total fixture bytes", tot)
