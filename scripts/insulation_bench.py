"""Config 5 (50 kb, chr2, 4864 bins, 10 Mb band: P = 948 036 pairs per cell) with the first consumer on the device: impute a block of
cells, reduce it to insulation scores / TAD boundary flags (hsg_insulation) and copy only those back.  Prints the time of the
two stages per block and the D2H bytes with and without the fused consumer.
Usage (GPU box): python scripts/insulation_bench.py [--cells 200] [--n-cells 400]"""
import argparse, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from higashi_b200 import engine as E, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=200)
ap.add_argument("--n-cells", type=int, default=400)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
ds = S.make_dataset("C5", shard=0, n_cells=a.n_cells)
sd, feats = S.random_state_dict(ds, ds.d, seed=0)
nbr, w = bench.knn_lists(ds, ds.k, 99)
eng = E.ScorerEngine(0)
eng.load_model(sd, ds.num, feats=feats, cell_feats=ds.cell_feats)
eng.set_contacts(ds.csr)
eng.set_neighbors(nbr, w)
P = eng.set_pairs(list(range(len(ds.bins))), ds.min_bin, ds.max_bin)
eng.prepare(E.DEFAULT_PRECISION, 0, E.ACT_SIGMOID)
n = int(ds.bins[0])
res = 50_000
w_bins, order = int(500_000 / res), int(500_000 / res / 2)
block = torch.empty((a.cells, P), dtype=torch.float32, device="cuda")
keep = np.ones(n, dtype=np.uint8); keep[::97] = 0
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
t_imp = t_ins = 0.0
for rep in range(a.reps + 1):
    c0 = (rep * a.cells) % max(1, ds.n_cells - a.cells + 1)
    ev[0].record()
    eng.impute_to_device(c0, c0 + a.cells, block)
    ev[1].record()
    score, border = eng.insulation(0, block, w_bins, order, keep, np.array([0, n - 1]))
    ev[2].record()
    torch.cuda.synchronize()
    if rep > 0:
        t_imp += ev[0].elapsed_time(ev[1]); t_ins += ev[1].elapsed_time(ev[2])
t_imp /= a.reps; t_ins /= a.reps
trip = a.cells * P
print("C5: %d cells x P=%d (n=%d bins, window %d bins): impute %.2f ms, insulation+boundaries %.2f ms per block -> %.3e triplets/s fused; "
      "D2H per block %.1f MB (scores) -> %.2f MB (profiles); boundaries per cell %.1f"
      % (a.cells, P, n, w_bins, t_imp, t_ins, trip / (t_imp + t_ins) * 1e3, trip * 4 / 1e6, a.cells * n * 5 / 1e6, border.sum() / a.cells))
