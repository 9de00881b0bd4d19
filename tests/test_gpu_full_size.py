"""Parity at BASELINE.json's full per-cell sizes through size-independent properties (the oracle cannot score 225 023 pairs x many
cells in seconds): the C2 genome (22 chromosomes, 2897 bins, P = 225 023 pairs per cell) with a reduced number of cells.

  * triplet enumeration: count = closed form, rows strictly increasing in generate_binpair order, every pair inside one chromosome;
  * determinism: the same call twice gives bit-identical scores;
  * batch-composition independence: imputing [0, n) in one call equals imputing it in unequal pieces, bit for bit (a cell's
    scores may not depend on which cells share its batch or its 128-triplet tiles);
  * precision: tensor-core path vs the fp32 path of the same library (itself pinned to the reference at 1e-5 on the fixtures)
    over every triplet of the block: max |diff| <= 1e-3, the tolerance BASELINE.json states for the 16-bit mode;
  * range: sigmoid scores lie in (0, 1) and are finite."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_16BIT = 1e-3


def test_c2_full_pair_set_properties():
    import bench
    from higashi_b200 import engine as E
    from higashi_b200 import synthetic as S
    n_cells = 96
    ds = S.make_dataset("C2", shard=0, n_cells=n_cells)
    sd, feats = S.random_state_dict(ds, ds.d, seed=0)
    nbr, w = bench.knn_lists(ds, ds.k, 99)
    chroms = list(range(len(ds.bins)))
    outs = {}
    for name, prec in (("fp32", E.PREC_FP32), ("tc16", E.PREC_TC16)):
        eng = E.ScorerEngine(0)
        eng.load_model(sd, ds.num, feats=feats, cell_feats=ds.cell_feats)
        eng.set_contacts(ds.csr)
        eng.set_neighbors(nbr, w)
        P = eng.set_pairs(chroms, ds.min_bin, ds.max_bin)
        assert P == sum(b * (b - 1) // 2 for b in ds.bins) == 225023
        if name == "fp32":
            off = 0
            for slot, b in enumerate(ds.bins):
                c = eng.coordinates(slot)
                assert c.shape == (b * (b - 1) // 2, 2) and c.min() >= 0 and c.max() < b
                assert np.all(c[:, 1] > c[:, 0])
                key = c[:, 0] * b + c[:, 1]
                assert np.all(np.diff(key) > 0)                       # generate_binpair order, no duplicates
                off += len(c)
            assert off == P
        eng.prepare(prec, 0, E.ACT_SIGMOID)
        n_use = 24 if name == "fp32" else n_cells                    # the fp32 CUDA-core path is ~20x slower
        full = eng.impute_to_host(0, n_use)
        again = eng.impute_to_host(0, n_use)
        assert np.array_equal(full, again)
        pieces = np.concatenate([eng.impute_to_host(a, b) for a, b in ((0, 5), (5, 6), (6, 19), (19, n_use))])
        assert np.array_equal(full, pieces)
        assert np.isfinite(full).all() and full.min() > 0.0 and full.max() < 1.0
        outs[name] = full
        eng.close()
    err = np.abs(outs["tc16"][:24] - outs["fp32"])
    print("C2 full pair set, 24 cells x 225023 pairs: max |tc16 - fp32| = %.3g, mean %.3g" % (err.max(), err.mean()))
    assert err.max() <= TOL_16BIT


@pytest.mark.parametrize("workload,n_cells,cells", [("C2", 48, [0, 17, 47]), ("C3", 24, [5]), ("C5", 24, [11])])
def test_full_size_cells_against_the_oracle(workload, n_cells, cells):
    """whole cells at BASELINE.json's per-cell sizes against the ORACLE (oracle.impute.impute_cells, pinned to
    the reference's goldens), not against another path of this library: C2 (P = 225 023 pairs per cell, 3 cells), one cell of C3
    (500 kb genome-wide, P = 899 k) and one of C5 (50 kb chr2, 10 Mb band, P = 948 k), in the fast tensor-core mode (1e-3),
    the high-accuracy mode (5e-4) and fp32 (1e-5)."""
    import torch
    import bench
    from higashi_b200 import engine as E
    from higashi_b200 import synthetic as S
    from oracle import impute as OI, scorer as OS
    ds = S.make_dataset(workload, shard=0, n_cells=n_cells)
    sd, feats = S.random_state_dict(ds, ds.d, seed=0)
    nbr, w = bench.knn_lists(ds, ds.k, 99)
    chroms = list(range(len(ds.bins)))
    tsd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
    with torch.no_grad():
        tables = OS.embed_tables(tsd, feats)
    mx = int(1e5) if ds.max_bin < 0 else ds.max_bin
    big, _, _, coords = OI.enumerate_pairs(ds.num_list, chroms, ds.min_bin, mx)
    ref = OI.impute_cells(tsd, tables, ds.cell_feats, ds.num_list, lambda c, cell: ds.csr.matrix(c, cell - 1), cells, chroms, big, 0,
                          nbr, w, "sigmoid")
    eng = E.ScorerEngine(0)
    eng.load_model(sd, ds.num, feats=feats, cell_feats=ds.cell_feats)
    eng.set_contacts(ds.csr)
    eng.set_neighbors(nbr, w)
    assert eng.set_pairs(chroms, ds.min_bin, ds.max_bin) == len(big)
    for j in chroms[:2]:
        assert np.array_equal(eng.coordinates(j), coords[j])
    for prec, tol, name in ((E.PREC_TC16, 1e-3, "tc16"), (E.PREC_TC16_HI, 5e-4, "tc16-hi"), (E.PREC_FP32, 1e-5, "fp32")):
        eng.prepare(prec, 0, E.ACT_SIGMOID)
        got = np.concatenate([eng.impute_to_host(c, c + 1) for c in cells])
        err = np.abs(got - ref)
        print("%s %s: %d cells x %d pairs vs oracle: max |err| = %.3g, mean %.3g" % (workload, name, len(cells), len(big), err.max(), err.mean()))
        assert err.max() <= tol, (workload, name, float(err.max()))
    eng.close()
