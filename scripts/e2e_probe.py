import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import bench
from higashi_b200 import engine as E, synthetic as S
class A: pass
args = A(); args.workload = "C2"; args.n_cells = 0
ds = S.make_dataset("C2", shard=0)
sd, feats = S.random_state_dict(ds, ds.d, seed=0)
nbr, w = bench.knn_lists(ds, ds.k, 99)
eng = E.ScorerEngine(0)
eng.load_model(sd, ds.num, feats=feats, cell_feats=ds.cell_feats)
P = eng.set_pairs(list(range(len(ds.bins))), ds.min_bin, ds.max_bin)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_rowptr, h_col, h_val = pin(ds.csr.row_ptr), pin(ds.csr.col), pin(ds.csr.val)
eng.ctx.set_contacts(h_rowptr.numpy(), h_col.numpy(), h_val.numpy())
eng.set_neighbors(nbr, w)
eng.prepare(E.PREC_TC16, 0, E.ACT_SIGMOID)
cps = 1000
h_out = torch.empty((cps, P), dtype=torch.float32).pin_memory()
out = torch.empty((cps, P), dtype=torch.float32, device="cuda")
nbr0 = np.ascontiguousarray(nbr - 1, dtype=np.int32)
T = {"dev": [], "host_out": [], "stage": [], "nbr": [], "commit": []}
for i in range(6):
    c0 = (i % 6) * cps
    torch.cuda.synchronize(); t = time.perf_counter()
    eng.impute_to_device(c0, c0 + cps, out, torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
    T["dev"].append(time.perf_counter() - t); t = time.perf_counter()
    eng.ctx.impute_cells_host_ptr(c0, c0 + cps, h_out.data_ptr())
    T["host_out"].append(time.perf_counter() - t); t = time.perf_counter()
    eng.ctx.stage_contacts(h_rowptr.numpy(), h_col.numpy(), h_val.numpy())
    T["stage"].append(time.perf_counter() - t); t = time.perf_counter()
    eng.ctx.set_neighbors(nbr0, w)
    T["nbr"].append(time.perf_counter() - t); t = time.perf_counter()
    eng.ctx.commit_contacts()
    T["commit"].append(time.perf_counter() - t)
for k, v in T.items(): print(k, ["%.1f" % (1e3 * x) for x in v])
