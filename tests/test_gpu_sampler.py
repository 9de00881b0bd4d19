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
