"""Mirror of the reference's imputation driver `higashi/Impute.py` on the B200 path.

`impute_process(config_path, model, name, mode, cell_start, cell_end, sparse_path, weighted_info=None)` keeps the
reference signature and semantics (Impute.py:123-302) and the argv CLI (Impute.py:319-343):
config JSON keys, `sparse_nondiag_adj_nbr_1.npy` (object array [C][n_cells] of scipy CSR), `weighted_info.npy`
([cell_neighbor_list, weight_dict]), output containers `{chrom}_{name}` with `coordinates` + `cell_{id}`.
What differs is only where the work runs: pair enumeration, per-cell neighbourhood preparation, node embedding
and the scorer are the K3 / K1 / K2 kernels of libhsg, driven per batch of cells instead of per cell."""
import os
import sys
import time

import numpy as np
import torch

from . import engine as E
from . import sink as S
from .Higashi_backend.utils import get_config, skip_start_end
from .synthetic import PackedCSR


def _neighbor_arrays(weighted_info, n_cells):
    """weighted_info.npy -> (nbr int64 [n_cells,k+1] 1-based self first, w float32) (Higashi_wrapper.py:1520-1545)."""
    cell_neighbor_list, weight_dict = weighted_info[0], weighted_info[1]
    k1 = len(cell_neighbor_list[1])
    nbr = np.zeros((n_cells, k1), dtype=np.int64)
    w = np.zeros((n_cells, k1), dtype=np.float32)
    for cell in range(1, n_cells + 1):
        lst = list(cell_neighbor_list[cell])
        if len(lst) != k1:
            raise ValueError("ragged neighbour lists are not supported")
        nbr[cell - 1] = lst
        w[cell - 1] = [weight_dict[(int(nb), cell)] for nb in lst]
    return nbr, w


def impute_process(config_path, model, name, mode, cell_start, cell_end, sparse_path, weighted_info=None,
                   precision=None, cells_per_call=512):
    if not torch.cuda.is_available():
        raise RuntimeError("impute_process runs on the B200 CUDA path only (no CPU fallback)")
    config = get_config(config_path)
    res = config["resolution"]
    impute_list, chrom_list, temp_dir = config["impute_list"], config["chrom_list"], config["temp_dir"]
    local_transfer_range = config["local_transfer_range"]
    impute_verbose = config.get("impute_verbose", 10)
    min_distance, max_distance = config["minimum_distance"], config["maximum_distance"]
    min_bin = 0 if min_distance < 0 else int(min_distance / res)
    max_bin = int(1e5) if max_distance < 0 else int(max_distance / res)
    min_bin_ = int(config["minimum_impute_distance"] / res) if "minimum_impute_distance" in config else min_bin
    if "maximum_impute_distance" in config:
        max_bin_ = int(1e5) if config["maximum_impute_distance"] < 0 else int(config["maximum_impute_distance"] / res)
    else:
        max_bin_ = max_bin

    num = S.read_num(temp_dir)
    n_cells = int(num[0])
    csr = PackedCSR.from_npy_cached(sparse_path)            # packed blob cache next to the pickled object array
    nbr = nbr_w = None
    if weighted_info is not None:
        nbr, nbr_w = _neighbor_arrays(np.load(weighted_info, allow_pickle=True), n_cells)

    model.eval()
    model.only_model = True
    eng = model.build_engine(csr, nbr, nbr_w)

    sel = [chrom_list.index(c) for c in impute_list]
    skip = []
    for c, j in zip(impute_list, sel):
        s_, e_ = skip_start_end(config, c)
        skip += [[j, int(a), int(b)] for a, b in zip(s_, e_)]
    eng.set_pairs(sel, min_bin_, max_bin_, skip)
    act = E.ACT_SIGMOID if mode == "classification" else E.ACT_SOFTPLUS
    eng.prepare(E.DEFAULT_PRECISION if precision is None else precision, local_transfer_range, act)
    print("total number of triplets to predict:", (eng.P, 3))

    sinks, offs = [], [0]
    for slot, chrom in enumerate(impute_list):
        sk = S.ChromSink(temp_dir, chrom, name)
        sk.create_dataset("coordinates", eng.coordinates(slot))
        sinks.append(sk)
        offs.append(offs[-1] + eng.sel_P[slot])

    start = time.time()
    # Two pinned blocks: the writer thread drains block k into the sinks (widen + copy + gzip) while the GPU fills block k + 1.
    # Block k + 2 reuses the buffer of block k, so it first waits for the TICKET of block k's write job (the queue bound of the
    # drain does not cover the job the writer is still running).  `impute_score_dtype: "float16"` in the config moves the scores
    # to the host as fp16 (half the D2H bytes, hsg_impute_cells_host_f16) and widens them here; the datasets stay float32.
    half = str(config.get("impute_score_dtype", "float32")) in ("float16", "fp16", "half") and eng.scorer_variant() == E.PREC_TC16
    rows = min(cells_per_call, max(1, cell_end - cell_start))
    bufs = [torch.empty((rows, max(1, eng.P)), dtype=torch.float16 if half else torch.float32).pin_memory() for _ in range(2)]
    tickets = [None, None]
    drain = S.AsyncDrain(depth=1)

    def write_block(out, c0, c1):
        for i in range(c0, c1):
            for slot in range(len(impute_list)):
                sinks[slot].create_dataset("cell_%d" % i, out[i - c0, offs[slot]:offs[slot + 1]].astype(np.float32), compress=True)

    count = 0
    for k, c0 in enumerate(range(cell_start, cell_end, cells_per_call)):
        c1 = min(cell_end, c0 + cells_per_call)
        buf = bufs[k % 2]
        if tickets[k % 2] is not None:
            tickets[k % 2].wait()                                        # block k - 2 has left this buffer
            drain._raise()
        eng.ctx.impute_cells_host_ptr(c0, c1, buf.data_ptr(), half=half)     # block k - 1 is being written meanwhile
        tickets[k % 2] = drain.submit(lambda out=buf.numpy()[: c1 - c0], a=c0, b=c1: write_block(out, a, b))
        count += c1 - c0
        if impute_verbose > 0:
            el = time.time() - start
            print("Imputing %s: %d of %d, takes %.2f s estimate %.2f s to finish" %
                  (name, count, cell_end - cell_start, el, el / count * (cell_end - cell_start - count)))
    drain.close()
    print("finish imputing, used %.2f s" % (time.time() - start))
    for sk in sinks:
        sk.close()
    model.train()
    model.only_model = False


if __name__ == "__main__":
    config_path, path, name, mode, cell_start, cell_end, sparse_path = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], \
        int(sys.argv[5]), int(sys.argv[6]), sys.argv[7]
    weighted_info = None if sys.argv[8] == "None" else sys.argv[8]
    gpu_id = sys.argv[9] if len(sys.argv) > 9 else "None"
    if gpu_id != "None":
        torch.cuda.set_device(int(gpu_id))
    from . import install_aliases
    install_aliases()
    model = torch.load(path, map_location="cpu", weights_only=False)
    impute_process(config_path, model, name, mode, cell_start, cell_end, sparse_path, weighted_info)
