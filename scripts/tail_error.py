"""Tail of the tensor-core path's score error at scale: fast (tcgen05, fp16 operands) vs the fp32 CUDA path of the same library on
`--cells` cells of a synthetic workload (both paths are parity-tested against the oracle; the fp32 path is within 1e-6 of it).
Usage (GPU box): python scripts/tail_error.py --workload C2 --cells 100"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from higashi_b200 import engine as E, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C2")
ap.add_argument("--cells", type=int, default=100)
ap.add_argument("--n-cells", type=int, default=600)
a = ap.parse_args()
ds = S.make_dataset(a.workload, shard=0, n_cells=a.n_cells)
sd, feats = S.random_state_dict(ds, ds.d, seed=0)
nbr, w = bench.knn_lists(ds, ds.k, 99)
outs = {}
for name, prec in (("fp32", E.PREC_FP32), ("tc16", E.PREC_TC16), ("hi", E.PREC_TC16_HI)):
    eng = E.ScorerEngine(0)
    eng.load_model(sd, ds.num, feats=feats, cell_feats=ds.cell_feats)
    eng.set_contacts(ds.csr)
    eng.set_neighbors(nbr, w)
    P = eng.set_pairs(list(range(len(ds.bins))), ds.min_bin, ds.max_bin)
    eng.prepare(prec, 0, E.ACT_SIGMOID)
    outs[name] = eng.impute_to_host(0, a.cells)
    eng.close()
for name in ("tc16", "hi"):
    err = np.abs(outs[name] - outs["fp32"])
    print("%s vs fp32 over %d triplets: max %.3e  p99.99 %.3e  mean %.3e" % (name, err.size, err.max(), np.quantile(err.ravel()[::7], 0.9999), err.mean()))
