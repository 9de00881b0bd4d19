"""GPU parity of the on-device insulation score / TAD boundary reduction (hsg_insulation) against the golden
vectors minted from the reference's own functions and against the CPU oracle on imputed scores."""
import numpy as np
import pytest
import torch

from tests._util import FixtureDataset, load

pytestmark = pytest.mark.gpu

TOL = 2e-6      # scores are stored as float32; the reduction itself runs in fp64


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_insulation_golden(tag):
    from higashi_b200.engine import ScorerEngine
    g = load("insulation.npz")
    n, w_ins, w_tad = [int(v) for v in g["%s/cfg" % tag]]
    coords, proba = g["%s/coords" % tag], g["%s/proba" % tag]
    band = int((coords[:, 1] - coords[:, 0]).max())
    eng = ScorerEngine(0)
    eng.ctx.set_dims(64, proba.shape[0], [11, n], 1)
    eng.num = [proba.shape[0], 11, n]
    P = eng.set_pairs([1], 1, band + 1)
    assert P == len(coords) and np.array_equal(eng.coordinates(0), coords)
    scores = torch.from_numpy(proba).cuda().contiguous()
    score, border = eng.insulation(0, scores, w_ins, w_tad, g["%s/keep" % tag], g["%s/discard" % tag])
    np.testing.assert_allclose(score, g["%s/score" % tag], atol=TOL, rtol=0)
    assert np.array_equal(border, g["%s/border" % tag])
    # no mask / no discarded rows
    from oracle import insulation as OX
    score, border = eng.insulation(0, scores, w_ins, w_tad)
    for c in range(proba.shape[0]):
        s_ref, b_ref = OX.cell_profile(coords, proba[c], n, np.ones(n), [], w_ins, w_tad)
        np.testing.assert_allclose(score[c], s_ref, atol=TOL, rtol=0)
        assert np.array_equal(border[c], b_ref)
    eng.close()


def test_insulation_after_imputation():
    """impute -> insulation without the scores leaving the device, against the oracle run on a host copy of the same scores."""
    from higashi_b200 import engine as E
    from higashi_b200.scTAD import gen_tad_scores
    from oracle import insulation as OX
    g = load("impute_k4_r0.npz")
    ds = FixtureDataset(g)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd/")}
    eng = E.ScorerEngine(0)
    eng.load_model(sd, ds.num, feats=ds.feats, cell_feats=ds.cell_feats)
    eng.set_contacts(ds.csr)
    eng.set_neighbors(g["knn/nbr"], g["knn/w"])
    chroms = list(range(len(ds.bins)))
    eng.set_pairs(chroms, 1, 12)
    eng.prepare(E.PREC_FP32, 0, E.ACT_SIGMOID)
    host = eng.impute_to_host(0, ds.n_cells)
    off = 0
    for slot, j in enumerate(chroms):
        n = ds.bins[j]
        keep = np.ones(n, dtype=np.uint8)
        keep[n // 3] = 0
        discard = np.array([1, n - 2])
        score, border, bulk = gen_tad_scores(eng, slot, 0, ds.n_cells, 6.0, 8.0, 1.0, keep, discard, cells_per_call=3)
        coords = eng.coordinates(slot)
        P = len(coords)
        np.testing.assert_allclose(bulk, host[:, off:off + P].astype(np.float64).sum(0), rtol=1e-6)
        for c in range(ds.n_cells):
            s_ref, b_ref = OX.cell_profile(coords, host[c, off:off + P], n, keep, discard, 6, 4)
            np.testing.assert_allclose(score[c], s_ref, atol=TOL, rtol=0)
            assert np.array_equal(border[c], b_ref)
        off += P
    # the full return tuple of the reference's gen_tad (single-cell part from the device, pseudo-bulk on the host)
    from higashi_b200.scTAD import gen_tad
    chrom, sc_score, sc_border, sc_idx, bulk_score, bulk_tad_b = gen_tad(eng, 0, "chr1", 0, ds.n_cells, 6.0, 8.0, 1.0, cells_per_call=7)
    n0 = ds.bins[chroms[0]]
    assert chrom == "chr1" and sc_score.shape == (ds.n_cells, n0) and sc_border.shape == (ds.n_cells, n0) and len(sc_idx) == ds.n_cells
    assert all(np.array_equal(np.flatnonzero(sc_border[i]), sc_idx[i]) for i in range(ds.n_cells))
    P0 = len(eng.coordinates(0))
    m = np.zeros((n0, n0)); np.add.at(m, (eng.coordinates(0)[:, 0], eng.coordinates(0)[:, 1]), host[:, :P0].astype(np.float64).sum(0))
    ref = OX.insulation_score(OX.sqrt_norm(m), 6)
    np.testing.assert_allclose(bulk_score, ref, rtol=1e-6, atol=1e-9)
    assert np.array_equal(bulk_tad_b, OX.call_tads(ref, 4))
    eng.close()


@pytest.mark.parametrize("n,min_bin,band,w_ins,order", [(70, 0, 15, 4, 2), (40, 1, 39, 60, 3), (33, 2, 9, 3, 1)])
def test_insulation_edge_cases(n, min_bin, band, w_ins, order):
    """Diagonal pairs in the list (min_bin = 0: m + m^T doubles them), a window wider than the chromosome, pairs starting at distance 2
    (a one-bin window over such a list makes every score exactly 1/2 in exact arithmetic -- the reference's boundary calls are then
    decided by its rounding noise, so that degenerate case is not compared); exact zeros in the scores."""
    from higashi_b200.engine import ScorerEngine
    from oracle import insulation as OX
    rng = np.random.Generator(np.random.PCG64(n))
    eng = ScorerEngine(0)
    eng.ctx.set_dims(64, 6, [n, 13], 1)
    eng.num = [6, n, 13]
    P = eng.set_pairs([0], min_bin, band + 1)
    coords = eng.coordinates(0)
    proba = rng.random((6, P)).astype(np.float32)
    proba[proba < 0.2] = 0.0
    proba[5] = 0.0
    keep = np.ones(n, dtype=np.uint8); keep[[2, n - 3]] = 0
    score, border = eng.insulation(0, torch.from_numpy(proba).cuda(), w_ins, order, keep, np.array([0]))
    for c in range(6):
        s_ref, b_ref = OX.cell_profile(coords, proba[c], n, keep, [0], w_ins, order)
        np.testing.assert_allclose(score[c], s_ref, atol=TOL, rtol=0)
        assert np.array_equal(border[c], b_ref)
    eng.close()
