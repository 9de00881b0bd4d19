"""Device-side mirror of the per-cell loop of the reference's single-cell TAD caller (`higashi/scTAD.py:gen_tad`, :60-121).

The reference reads every imputed `cell_%d` vector back from HDF5, scatters it into an n x n matrix, normalises it and computes
the insulation profile on one CPU core.  Here the scores never leave the GPU: `hsg_impute_cells` writes a block of cells and
`hsg_insulation` reduces it to the insulation score / boundary flags (Config 5 exists for this use)."""
import numpy as np
import torch


def gen_tad_scores(eng, slot, cell_start, cell_end, window_ins, window_tad, res, keep=None, discard_rows=None, cells_per_call=128):
    """eng: a prepared ScorerEngine (set_pairs / prepare done).  Returns (sc_score [cells, n] float32, sc_border [cells, n] uint8,
    bulk1 [P_slot] float64 = sum of the imputed vectors over the cells, scTAD.py:100).

    window_bins = int(window_ins / res) (Higashi_TAD.py:8), tad_order = int(window_tad / res / 2) (Higashi_TAD.py:23)."""
    if not torch.cuda.is_available():
        raise RuntimeError("gen_tad_scores runs on the B200 CUDA path only (no CPU fallback)")
    w_bins, order = int(window_ins / res), int(window_tad / res / 2)
    dev = torch.device("cuda", eng.device)
    off = int(sum(eng.sel_P[:slot]))
    P_slot = int(eng.sel_P[slot])
    block = torch.empty((min(cells_per_call, max(1, cell_end - cell_start)), max(1, eng.P)), dtype=torch.float32, device=dev)
    bulk = torch.zeros(P_slot, dtype=torch.float64, device=dev)
    scores, borders = [], []
    for c0 in range(cell_start, cell_end, cells_per_call):
        c1 = min(cell_end, c0 + cells_per_call)
        view = block[: c1 - c0]
        eng.impute_to_device(c0, c1, view)
        torch.cuda.synchronize(dev)
        s, b = eng.insulation(slot, view, w_bins, order, keep, discard_rows)
        bulk += view[:, off:off + P_slot].sum(dim=0, dtype=torch.float64)
        scores.append(s)
        borders.append(b)
    return np.concatenate(scores), np.concatenate(borders), bulk.cpu().numpy()
