"""The Python host mirror of the reference interface (Higashi_backend.Modules, Impute.impute_process):
state_dict compatibility and pickling on CPU; the end-to-end file-level imputation on the GPU."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from tests._util import FixtureDataset, load


def build_mirror_model(ds, d, sd=None):
    from higashi_b200.Higashi_backend import Modules as M
    feats = [ds.cell_embed_feats] + list(ds.bin_feats)
    emb = M.MultipleEmbedding(feats, d, False, ds.num_list, feats)
    model = M.Hyper_SAGNN(n_head=8, d_model=d, d_k=16, d_v=16, diag_mask=True, bottle_neck=d,
                          attribute_dict=ds.attribute_dict, cell_feats=ds.cell_feats, encoder_dynamic_nn=emb,
                          encoder_static_nn=emb, chrom_num=len(ds.bins))
    n_nodes = int(ds.num_list[-1])
    enc = M.GraphSageEncoder_with_weights(features=emb, linear_features=emb, feature_dim=d, embed_dim=d, num_sample=8,
                                          gcn=False, num_list=ds.num_list, transfer_range=0,
                                          start_end_dict=np.zeros((n_nodes + 1, 2), dtype="int"), pass_pseudo_id=False,
                                          remove=True, pass_remove=False)
    model.encode1.dynamic_nn = enc
    if sd is not None:
        missing, unexpected = model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
        assert not unexpected, unexpected
    return model


def test_state_dict_is_reference_compatible():
    """Every tensor of a reference state_dict (golden fixture) has a same-named, same-shaped slot in the mirror."""
    g = load("impute_k4_r0.npz")
    ds = FixtureDataset(g)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd/")}
    model = build_mirror_model(ds, 64, sd)
    own = model.state_dict()
    for k, v in sd.items():
        assert k in own, k
        assert tuple(own[k].shape) == tuple(v.shape), k
        np.testing.assert_array_equal(own[k].numpy(), v)


def test_model_pickles_without_engine(tmp_path):
    g = load("impute_k4_r0.npz")
    ds = FixtureDataset(g)
    model = build_mirror_model(ds, 64)
    p = tmp_path / "model.pkl"
    torch.save(model, p)
    m2 = torch.load(p, weights_only=False)
    assert type(m2).__name__ == "Hyper_SAGNN" and m2._engine is None
    assert m2.encode1.dynamic_nn.forward.__name__ == "forward_on_hook"


def test_reference_module_aliases():
    import sys
    import higashi_b200
    higashi_b200.install_aliases()
    from higashi_b200.Higashi_backend import Modules as M
    assert sys.modules["Higashi_backend.Modules"] is M
    assert pickle.loads(pickle.dumps(M.Hyper_SAGNN)) is M.Hyper_SAGNN


def test_data_generator_matches_reference_semantics():
    from higashi_b200.Higashi_backend.Modules import DataGenerator
    rng = np.random.default_rng(0)
    edges = [np.sort(rng.integers(1, 50, size=(30, 3)), axis=-1) for _ in range(2)]
    gen = DataGenerator(edges, [np.zeros(30, "int8"), np.ones(30, "int8")], [np.ones(30), np.ones(30)], 8, True,
                        num_list=[10, 30, 50], k=1)
    np.random.seed(1)
    e, c, w, chroms = gen.next_iter()
    assert e.shape == (8, 3) and len(c) == 8 and len(chroms) == 1 and (c == chroms[0]).all()
    gen.filter_edges(2, 20)
    for ed in gen.edges:
        dist = ed[:, 2] - ed[:, 1]
        assert ((dist > 2) & (dist < 20)).all()


def test_forward_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = load("impute_k4_r0.npz")
    model = build_mirror_model(FixtureDataset(g), 64)
    with pytest.raises(RuntimeError):
        model.predict(np.ones((3, 3), dtype=np.int64), np.zeros(3))


# ------------------------------------------------------------------------------------------ GPU: file-level run
def _write_run_dir(tmp, g, ds, chrom_list):
    cfg = dict(resolution=1000000, impute_list=chrom_list, chrom_list=chrom_list, temp_dir=str(tmp), minimum_distance=1000000,
               maximum_distance=-1, local_transfer_range=int(g["cfg/ltr"]), embedding_name="emb", impute_verbose=0)
    mn, mx = [int(v) for v in g["cfg/band"]]
    if mx > 0:
        cfg["minimum_impute_distance"] = mn * 1000000
        cfg["maximum_impute_distance"] = mx * 1000000
    if len(g["cfg/skip"]):
        # a cytoband file that reproduces the fixture's skip intervals: bins [a,b) <- acen [ (a)*res+1e5 , (b)*res-1e5 ]
        with open(tmp / "cyto.txt", "w") as f:
            for j, a, b in g["cfg/skip"]:
                f.write("%s\t%d\t%d\tx\tacen\n" % (chrom_list[int(j)], int(a) * 1000000 + 100000, int(b) * 1000000 - 100000))
        cfg["cytoband_path"] = str(tmp / "cyto.txt")
    json.dump(cfg, open(tmp / "config.json", "w"))
    np.savez(tmp / "node_feats.npz", num=ds.num)
    np.save(tmp / "sparse_nondiag_adj_nbr_1.npy", ds.csr.to_object_array(), allow_pickle=True)
    winfo = None
    if int(g["cfg/k"]) > 0:
        nl = [[]] + [list(r) for r in g["knn/nbr"]]
        wd = {}
        for i in range(1, len(nl)):
            for c, w in zip(nl[i], g["knn/w"][i - 1]):
                wd[(int(c), i)] = np.float32(w)
        np.save(tmp / "weighted_info.npy", np.array([np.array(nl, dtype=object), wd], dtype=object), allow_pickle=True)
        winfo = str(tmp / "weighted_info.npy")
    return str(tmp / "config.json"), str(tmp / "sparse_nondiag_adj_nbr_1.npy"), winfo


@pytest.mark.gpu
@pytest.mark.parametrize("tag,prec", [("k4_r0", "fp32"), ("k4_r0", "tc16"), ("k0_r1_band", "fp32"), ("k4_r1", "tc16")])
def test_impute_process_files(tmp_path, tag, prec):
    """impute_process(config, model, name, mode, cell_start, cell_end, sparse_path, weighted_info) end to end."""
    from higashi_b200 import engine as E
    from higashi_b200 import sink
    from higashi_b200.Impute import impute_process
    g = load("impute_%s.npz" % tag)
    ds = FixtureDataset(g)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd/")}
    model = build_mirror_model(ds, 64, sd)
    chrom_list = ["chr%d" % (j + 1) for j in range(len(ds.bins))]
    cfg_path, sparse_path, winfo = _write_run_dir(tmp_path, g, ds, chrom_list)
    c0, c1 = [int(v) for v in g["cfg/cell_range"]]
    name = "emb_nbr_%d_impute" % int(g["cfg/k"])
    impute_process(cfg_path, model, name, str(g["cfg/mode"]), c0, c1, sparse_path, winfo,
                   precision=E.PREC_FP32 if prec == "fp32" else E.PREC_TC16)
    tol = 1e-5 if prec == "fp32" else 1e-3
    for j, chrom in enumerate(chrom_list):
        f = sink.read_container(str(tmp_path), chrom, name)
        ref_c = g["out/%s/coordinates" % chrom]
        assert f["coordinates"].dtype == np.int64 and np.array_equal(f["coordinates"], ref_c)
        for n, cell in enumerate(range(c0, c1)):
            v = f["cell_%d" % cell]
            assert v.dtype == np.float32 and v.shape == (len(ref_c),)
            np.testing.assert_allclose(v, g["out/%s/cells" % chrom][n], atol=tol, rtol=0)
    assert model.training and model.only_model is False        # state restored like Impute.py:288-292


@pytest.mark.gpu
def test_predict_api_like_reference():
    """model.encode1.dynamic_nn.fix_cell2(cell, ...) then model.predict(triplets, chrom, activation=torch.sigmoid)."""
    g = load("scorer_d64.npz")
    ds = FixtureDataset(g)
    sd = {k[len("stage2/sd/"):]: v for k, v in g.items() if k.startswith("stage2/sd/")}
    model = build_mirror_model(ds, 64, sd)
    model.eval()
    model.only_model = True
    model.build_engine(ds.csr)
    pairs = g["impute/pairs"]
    for n, cell in enumerate(g["impute/cells"]):
        x = np.concatenate([np.full((len(pairs), 1), int(cell), dtype=np.int64), pairs], axis=-1)
        model.encode1.dynamic_nn.fix_cell2(int(cell), None, None, 0, route_nn_list=None)
        out = model.predict(x, np.zeros(len(x), dtype=int), verbose=False, batch_size=int(1e5), activation=torch.sigmoid)
        assert out.shape == (len(x), 1) and out.dtype == np.float32
        np.testing.assert_allclose(out.reshape(-1), g["impute/sigmoid"][n], atol=1e-5, rtol=0)


@pytest.mark.gpu
def test_predict_honours_local_transfer_range_and_survives_pickle():
    """the drop-in loop of Impute.py:253-277 (fix_cell2(cell, ..., local_transfer_range) then predict) must
    score with the requested local_transfer_range (golden fixture with range 1), and a model that went through torch.save /
    torch.load or a rebuilt engine must prepare the new engine again instead of failing with "call hsg_prepare first"."""
    import io
    g = load("impute_k0_r1_band.npz")
    ds = FixtureDataset(g)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd/")}
    model = build_mirror_model(ds, 64, sd)
    model.eval()
    model.only_model = True
    ltr = int(g["cfg/ltr"])
    assert ltr == 1
    c0, c1 = [int(v) for v in g["cfg/cell_range"]]

    def check(m):
        m.build_engine(ds.csr)
        for cell in (c0, c1 - 1):
            m.encode1.dynamic_nn.fix_cell2(cell + 1, None, None, ltr, route_nn_list=None)
            for j in range(len(ds.bins)):
                coords = g["out/chr%d/coordinates" % (j + 1)]
                base = int(ds.num_list[j]) + 1
                x = np.concatenate([np.full((len(coords), 1), cell + 1, dtype=np.int64), coords + base], axis=-1)
                out = m.predict(x, np.full(len(x), j, dtype=int), activation=torch.sigmoid)
                np.testing.assert_allclose(out.reshape(-1), g["out/chr%d/cells" % (j + 1)][cell - c0], atol=1e-5, rtol=0)
    check(model)
    check(model)                       # rebuilt engine: prepared again
    buf = io.BytesIO()
    torch.save(model, buf)
    buf.seek(0)
    clone = torch.load(buf, map_location="cpu", weights_only=False)
    assert getattr(clone, "_prepared", None) is None and clone._engine is None
    check(clone)


def test_packed_contact_blob_round_trip(tmp_path):
    """Stable binary layout of the packed contact CSR (8f rank 3): save -> memmap load -> identical arrays, and the object-array
    route of the reference (sparse_nondiag_adj_nbr_1.npy) packs to the same blob."""
    from higashi_b200 import synthetic as S
    ds = S.make_dataset("C1", n_cells=7, bins=[31, 12], seed=2)
    path = str(tmp_path / "contacts.hsgcsr")
    ds.csr.save_blob(path)
    for mm in (True, False):
        got = S.PackedCSR.load_blob(path, mmap=mm)
        assert got.n_cells == ds.csr.n_cells and np.array_equal(got.bins, ds.csr.bins) and np.array_equal(got.bin_off, ds.csr.bin_off)
        assert np.array_equal(got.row_ptr, ds.csr.row_ptr) and np.array_equal(got.col, ds.csr.col) and np.array_equal(got.val, ds.csr.val)
    again = S.PackedCSR.from_object_array(ds.csr.to_object_array())
    path2 = str(tmp_path / "contacts2.hsgcsr")
    again.save_blob(path2)
    assert open(path, "rb").read() == open(path2, "rb").read()
    with open(path, "r+b") as f:
        f.write(b"XXXX")
    import pytest
    with pytest.raises(ValueError):
        S.PackedCSR.load_blob(path)


def test_async_drain_orders_jobs_and_propagates_errors():
    """The background writer of the output sink (8f rank 2): jobs run in submission order, at most depth jobs are pending, and an
    exception raised by a job surfaces in the producer."""
    import time
    import pytest
    from higashi_b200.sink import AsyncDrain
    seen = []
    d = AsyncDrain(depth=1)
    for i in range(6):
        d.submit(lambda i=i: (time.sleep(0.01), seen.append(i)))
    d.close()
    assert seen == list(range(6))
    d = AsyncDrain(depth=1)
    d.submit(lambda: (_ for _ in ()).throw(ValueError("disk full")))
    with pytest.raises(ValueError):
        d.close()


def test_recycled_score_buffers_wait_for_their_writer():
    """the producer of Impute.impute_process recycles two pinned blocks.  A slow writer (gzip) must never see
    a block overwritten under it: block k + 2 may only be filled after the ticket of block k's write job is set.  Same protocol
    as Impute.py, CPU only, 6 blocks, writer 20x slower than the producer."""
    import time
    from higashi_b200.sink import AsyncDrain
    bufs = [np.zeros(64, dtype=np.float32) for _ in range(2)]
    tickets = [None, None]
    written = {}
    drain = AsyncDrain(depth=1)

    def write(view, k):
        time.sleep(0.02)                      # slow sink
        written[k] = view.copy()              # lazy copy, as write_block does per cell

    for k in range(6):
        if tickets[k % 2] is not None:
            tickets[k % 2].wait()
        bufs[k % 2][:] = float(k)             # the GPU + D2H refill of this block
        tickets[k % 2] = drain.submit(lambda v=bufs[k % 2], k=k: write(v, k))
    drain.close()
    assert sorted(written) == list(range(6))
    for k in range(6):
        assert np.all(written[k] == float(k)), (k, written[k][:4])


# ------------------------------------------------------------------ generation-4 tile table (host logic of the fast scorer)
def _tile_table(pairs, allow_dense):
    import ctypes as C
    from higashi_b200 import _lib
    lib = _lib.load()
    pairs = np.ascontiguousarray(pairs, dtype=np.int32)
    out = np.zeros((len(pairs) + 1, 8), dtype=np.int32)
    n = lib.hsg_debug_tile_table(pairs.ctypes.data_as(C.POINTER(C.c_int32)), len(pairs), int(allow_dense),
                                 out.ctypes.data_as(C.POINTER(C.c_int32)), len(out))
    assert n >= 0
    return out[:n]


@pytest.mark.parametrize("bins,band", [([120, 40], -1), ([250, 97, 30], -1), ([300], 40), ([5, 3, 129], -1)])
def test_tile_table_covers_the_pair_list(bins, band):
    """score_umma4.cu `build_tiles4` (pure host code): tiles are consecutive runs of the pair list in generate_binpair order, cover
    it exactly once, hold at most 128 pairs; a plain tile has at most three bin_i segments with the recorded boundaries and bins;
    dense tiles (the short rows at the end of a chromosome) appear only where the three-segment prefix would leave more than 16 rows
    idle, hold 128 pairs unless the list ends, and make the table strictly shorter."""
    from oracle import impute as OI
    num_list = np.concatenate([[0], np.cumsum(bins)]) + 10                      # bins are numbered after 10 cells
    big, _, _, _ = OI.enumerate_pairs(num_list, list(range(len(bins))), 1, int(1e5) if band < 0 else band)
    pairs = np.asarray(big)[:, -2:].astype(np.int32)
    P = len(pairs)
    plain, dense = _tile_table(pairs, False), _tile_table(pairs, True)
    for tab, allow in ((plain, False), (dense, True)):
        p0 = 0
        for (s0, cnt, s1, s2, b0, b1, b2, flag) in tab:
            assert s0 == p0 and 1 <= cnt <= 128
            bi = pairs[p0:p0 + cnt, 0]
            starts = np.flatnonzero(np.r_[True, bi[1:] != bi[:-1]])
            if flag:
                assert allow and s1 == 0 and s2 == 0 and (cnt == 128 or p0 + cnt == P) and len(starts) > 3
                assert starts[3] < 112                                          # the three-segment prefix was short
            else:
                assert len(starts) <= 3
                want = list(starts[1:]) + [128] * (3 - len(starts))
                assert [s1, s2] == want
                segb = list(bi[starts]) + [bi[0]] * (3 - len(starts))
                assert [b0, b1, b2] == segb
                if allow and p0 + cnt < P and cnt < 128:
                    assert cnt >= 112                                           # a shorter prefix would have become a dense tile
            p0 += cnt
        assert p0 == P
    assert len(dense) <= len(plain)
    if len(bins) > 1 or band < 0:
        assert len(dense) < len(plain)
