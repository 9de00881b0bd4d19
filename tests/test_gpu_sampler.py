"""Training batch producer on the device (hsg_sample_batch): negatives, labels, weights and the neighbour CSR of a minibatch.
The reference producer is unseeded Python (generate_negative_cpu / one_thread_generate_neg, Higashi_wrapper.py:95-333), so the
draws are checked for their invariants and statistics, everything deterministic is checked exactly against the host
restatements (batches.neighbor_csr, NegativeSampler._nonzero_near), and the forward / backward on the device batch must equal
the host-fed path on the same arrays."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MAX_BIN = 60


def _setup(k=4, seed=11, n_cells=60, bins=(150, 90), density=0.03):
    from higashi_b200 import batches, synthetic as S
    from higashi_b200.train_engine import TrainEngine
    from oracle import prep as OP
    ds = S.make_dataset("C1", n_cells=n_cells, bins=list(bins), seed=seed, density=density)
    sd, feats = S.random_state_dict(ds, 64, seed=3)
    eng = TrainEngine(sd, ds.num, feats, ds.cell_feats, device=0)
    nbr = nbr_w = None
    if k > 0:
        rng = np.random.Generator(np.random.PCG64(9))
        nbr, nbr_w = OP.cell_neighbors(rng.standard_normal((ds.n_cells, 16)).astype(np.float32), k + 1)
        nbr_w = nbr_w.astype(np.float32)
    eng.set_graph(ds.csr, nbr, nbr_w)
    rng = np.random.default_rng(5)
    pos_all, _, _ = batches.synthetic_batch(ds.csr, 0, ds.num_list, 200, 0, rng)
    pos = pos_all[:200]
    pos = pos[(pos[:, 2] - pos[:, 1] > 1) & (pos[:, 2] - pos[:, 1] < MAX_BIN)]
    pos_w = rng.uniform(0.5, 2.0, size=len(pos)).astype(np.float32)
    return ds, eng, nbr, nbr_w, pos, pos_w


def _rows(indptr, rowcnt, cols, vals):
    return [(cols[indptr[r]:indptr[r] + rowcnt[r]], vals[indptr[r]:indptr[r] + rowcnt[r]]) for r in range(len(rowcnt))]


@pytest.mark.parametrize("k", [0, 4])
def test_neighbour_rows_exact_without_removal(k):
    from higashi_b200 import batches
    ds, eng, nbr, nbr_w, pos, pos_w = _setup(k=k)
    B, nnz = eng.sample_batch(pos, 0, 3, 0.5, MAX_BIN, pos_w=pos_w, training=False, seed=1)
    x, y, w, indptr, rowcnt, cols, vals = eng.fetch_sampled(B, nnz)
    Bp = len(pos)
    assert np.array_equal(x[:Bp], pos) and np.all(y[:Bp] == 1) and np.allclose(w[:Bp], pos_w)
    assert Bp < B <= 4 * Bp and np.all(y[Bp:] == 0) and np.all(w[Bp:] == 1.0)
    ip, cc, vv = batches.neighbor_csr(ds.csr, x, 0, ds.num_list, training=False, nbr=nbr, nbr_w=nbr_w)
    got = _rows(indptr, rowcnt, cols, vals)
    assert np.array_equal(rowcnt, np.diff(ip))
    for r, (gc, gv) in enumerate(got):
        assert np.array_equal(gc, cc[ip[r]:ip[r + 1]]), r
        np.testing.assert_allclose(gv, vv[ip[r]:ip[r + 1]], rtol=2e-6, atol=1e-7)
    eng.close()


def test_negatives_respect_the_acceptance_rule():
    from higashi_b200.Higashi_wrapper import NegativeSampler
    ds, eng, nbr, nbr_w, pos, pos_w = _setup(k=0, n_cells=80)
    neg_num, ratio = 5, 0.6
    B, nnz = eng.sample_batch(pos, 0, neg_num, ratio, MAX_BIN, training=False, seed=77)
    x, y, w, *_ = eng.fetch_sampled(B, nnz)
    Bp = len(pos)
    neg = x[Bp:]
    base, n = int(ds.num_list[0]) + 1, int(ds.bins[0])
    assert len(neg) >= 0.95 * Bp * neg_num                                   # almost every draw succeeds within 50 trials
    assert np.all((neg[:, 0] >= 1) & (neg[:, 0] <= ds.n_cells))
    assert np.all((neg[:, 1] >= base) & (neg[:, 2] < base + n))
    dist = neg[:, 2] - neg[:, 1]
    assert np.all((dist > 1) & (dist < MAX_BIN))
    ns = NegativeSampler(ds.csr, ds.num_list, MAX_BIN)
    assert not ns._nonzero_near(0, neg[:, 0] - 1, neg[:, 1] - base, neg[:, 2] - base).any()     # never on / next to a contact
    if len(neg) == Bp * neg_num:                                             # origin of every negative is known: q // neg_num
        src = np.repeat(pos, neg_num, axis=0)
        shared = (neg[:, 0] == src[:, 0]).astype(int) + (neg[:, 1] == src[:, 1]) + (neg[:, 2] == src[:, 2]) + \
            (neg[:, 1] == src[:, 2]) + (neg[:, 2] == src[:, 1])              # the bins are re-sorted after the change
        hard = shared == 2                                                   # exactly one node changed
        assert abs(hard.mean() - ratio) < 0.12
    # same seed -> same batch, another seed -> another batch
    B2, nnz2 = eng.sample_batch(pos, 0, neg_num, ratio, MAX_BIN, training=False, seed=77)
    x2 = eng.fetch_sampled(B2, nnz2)[0]
    assert B2 == B and np.array_equal(x2, x)
    B3, nnz3 = eng.sample_batch(pos, 0, neg_num, ratio, MAX_BIN, training=False, seed=78)
    assert not np.array_equal(eng.fetch_sampled(B3, nnz3)[0][Bp:Bp + 50], x[Bp:Bp + 50])
    eng.close()


def test_partner_removal_statistics():
    from higashi_b200 import batches
    ds, eng, nbr, nbr_w, pos, pos_w = _setup(k=4, density=0.06)
    B, nnz = eng.sample_batch(pos, 0, 0, 0.5, MAX_BIN, training=True, seed=5)       # positives only: the partner is always a contact
    x, _, _, indptr, rowcnt, cols, vals = eng.fetch_sampled(B, nnz)
    ip, cc, vv = batches.neighbor_csr(ds.csr, x, 0, ds.num_list, training=False, nbr=nbr, nbr_w=nbr_w)
    base = int(ds.num_list[0]) + 1
    partner = x[:, [2, 1]].reshape(-1) - base
    removed = eligible = 0
    for r, (gc, gv) in enumerate(_rows(indptr, rowcnt, cols, vals)):
        full = cc[ip[r]:ip[r + 1]]
        if len(full) > 1 and partner[r] in full:
            eligible += 1
            if len(gc) == len(full) - 1:
                removed += 1
                assert np.array_equal(gc, full[full != partner[r]])
                assert abs(gv.sum() - 1.0) < 1e-5                            # renormalised after the removal
            else:
                assert np.array_equal(gc, full)
        else:
            assert np.array_equal(gc, full)
    assert eligible > 100 and abs(removed / eligible - 0.4) < 0.1
    eng.close()


def test_forward_backward_on_device_batch_equals_host_fed_batch():
    ds, eng, nbr, nbr_w, pos, pos_w = _setup(k=4)
    B, nnz = eng.sample_batch(pos, 0, 4, 0.5, MAX_BIN, pos_w=pos_w, training=True, seed=9)
    x, y, w, indptr, rowcnt, cols, vals = eng.fetch_sampled(B, nnz)
    mean_d, loss_d = eng.forward(None, 0, sampled=(B, nnz), stage=2)
    eng.zero_grad(); eng.backward(); torch.cuda.synchronize()
    g_dev = eng.grads.clone()
    # compact CSR of the same rows through the host-fed path
    ip = np.zeros(2 * B + 1, np.int32); np.cumsum(rowcnt, out=ip[1:])
    cc = np.concatenate([cols[indptr[r]:indptr[r] + rowcnt[r]] for r in range(2 * B)]).astype(np.int32)
    vv = np.concatenate([vals[indptr[r]:indptr[r] + rowcnt[r]] for r in range(2 * B)]).astype(np.float32)
    mean_h, loss_h = eng.forward(x, 0, (ip, cc, vv), y=y, w=w, stage=2)
    eng.zero_grad(); eng.backward(); torch.cuda.synchronize()
    np.testing.assert_allclose(mean_d, mean_h, atol=1e-6, rtol=0)
    assert abs(loss_d - loss_h) < 1e-6
    scale = float(eng.grads.abs().max())
    assert float((g_dev - eng.grads).abs().max()) <= 1e-5 * max(scale, 1.0)
    eng.close()


@pytest.mark.parametrize("tag", ["plain_eval", "knn_eval", "plain_train", "knn_train"])
def test_device_neighbour_lists_match_the_reference_golden(tag):
    """the device neighbour CSR against the UNMODIFIED reference (tests/golden/neighs.npz was minted by
    one_thread_generate_neg + to_neighs_to_mask, Higashi_wrapper.py:236-333, 171-221), not against the repo's own host restatement.
    The golden hyperedges (positives and the reference's negatives alike) go in as the batch, neg_num = 0; the `*_train` cases
    replay the reference's stored uniform draws of the 40 % partner removal through hsg_sample_set_removal_draws."""
    from higashi_b200 import synthetic as S
    from higashi_b200.train_engine import TrainEngine
    from tests._util import FixtureDataset, load
    g = load("neighs.npz")
    ds = FixtureDataset(g)
    sd, _ = S.random_state_dict(ds, 64, seed=3)
    eng = TrainEngine(sd, ds.num, ds.feats, ds.cell_feats, device=0)
    knn = tag.startswith("knn")
    eng.set_graph(ds.csr, g["knn/nbr"] if knn else None, g["knn/w"].astype(np.float32) if knn else None)
    x, x_chrom = g[tag + "/x"], g[tag + "/x_chrom"]
    training = bool(g[tag + "/training"])
    draws = g[tag + "/remove_draw"].astype(np.float32)
    ind, val, uniq = g[tag + "/indices"], g[tag + "/values"], g[tag + "/unique"]
    # golden rows: token row 2r + t of hyperedge r, columns = index into `unique` (global node ids, first-seen order)
    gold = {}
    for row, cidx, v in zip(ind[0], ind[1], val):
        gold.setdefault(int(row), []).append((int(uniq[cidx]), float(v)))
    checked = 0
    for chrom in np.unique(x_chrom):
        sel = np.nonzero(x_chrom == chrom)[0]
        xs = np.ascontiguousarray(x[sel])
        assert np.all(xs[:, 1] <= xs[:, 2])
        if training:
            eng.set_removal_draws(np.stack([draws[2 * sel], draws[2 * sel + 1]], axis=1).reshape(-1))
        B, nnz = eng.sample_batch(xs, int(chrom), 0, 0.5, 10 ** 6, training=training, seed=1)
        assert B == len(xs)
        xb, _, _, indptr, rowcnt, cols, vals = eng.fetch_sampled(B, nnz)
        assert np.array_equal(xb, xs)
        base = int(ds.num_list[chrom]) + 1
        for i, r in enumerate(sel):
            for t in range(2):
                lo, n = indptr[2 * i + t], rowcnt[2 * i + t]
                want = sorted(gold.get(2 * int(r) + t, []))
                assert n == len(want), (tag, int(r), t)
                assert np.array_equal(cols[lo:lo + n] + base, np.array([w_[0] for w_ in want], dtype=np.int64)), (tag, int(r), t)
                np.testing.assert_allclose(vals[lo:lo + n], np.array([w_[1] for w_ in want], dtype=np.float32), rtol=2e-6, atol=1e-7)
                checked += n
    assert checked == len(val)
    eng.set_removal_draws(None)
    eng.close()
