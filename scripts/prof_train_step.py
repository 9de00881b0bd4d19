import cProfile, pstats, sys, os, numpy as np, torch, time
sys.path.insert(0, ".")
from higashi_b200 import batches, parallel, synthetic as S
from higashi_b200.train_engine import TrainEngine
ds = S.make_dataset("C4", n_cells=300, seed=4000, density=0.02)
sd, feats = S.random_state_dict(ds, ds.d, seed=0)
eng = TrainEngine(sd, ds.num, feats, ds.cell_feats, device=0)
tr = parallel.DataParallelTrainer(eng)
rng = np.random.default_rng(1)
x, y, w = batches.synthetic_batch(ds.csr, 0, ds.num_list, 512, 5, rng)
nb = batches.neighbor_csr(ds.csr, x, 0, ds.num_list)
for i in range(10): tr.step(x, 0, nb, y, w, want_loss=False)
torch.cuda.synchronize()
# GPU-only time: events around 50 steps
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for i in range(100): tr.step(x, 0, nb, y, w, want_loss=False)
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print("wall per step %.3f ms, event per step %.3f ms" % ((t1 - t0) * 10, e0.elapsed_time(e1) / 100))
pr = cProfile.Profile(); pr.enable()
for i in range(100): tr.step(x, 0, nb, y, w, want_loss=False)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
