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


def _window_sums(m, w):
    """num[i] = sum(m[i-w:i, i+1:min(n-1, i+w+1)]), den[i] = sum(m[i-w:i+w+1, i-w:i+w+1]) for every bin i, from one 2-D prefix sum
    (host side, once per chromosome: the pseudo-bulk profile of scTAD.py:113-119)."""
    n = m.shape[0]
    c = np.zeros((n + 1, n + 1), dtype=np.float64)
    c[1:, 1:] = m.cumsum(0).cumsum(1)

    def block(r0, r1, c0, c1):          # sum of m[r0:r1, c0:c1] for index arrays (empty ranges give 0)
        r1 = np.maximum(r1, r0); c1 = np.maximum(c1, c0)
        return c[r1, c1] - c[r0, c1] - c[r1, c0] + c[r0, c0]

    i = np.arange(n)
    lo, hi = np.maximum(0, i - w), np.minimum(n, i + w + 1)
    num = block(lo, i, np.minimum(i + 1, n), np.minimum(n - 1, i + w + 1))
    den = block(lo, hi, lo, hi)
    return num, den


def bulk_profile(coordinates, bulk, n, window_ins, window_tad, res, keep=None, discard_rows=None):
    """Pseudo-bulk insulation score and TAD boundaries exactly as scTAD.gen_tad builds them (scTAD.py:113-119): the summed
    imputed vector is scattered into the UPPER triangle only (the reference does not symmetrise the bulk), masked, sqrt_norm'ed
    (Higashi_analysis.py:581-588), then insulation_score / call_tads (Higashi_TAD.py:7-27).  Returns (bulk_score float64 [n],
    bulk_tad_b int64 [...])."""
    w, order = int(window_ins / res), int(window_tad / res / 2)
    m = np.zeros((n, n), dtype=np.float64)
    np.add.at(m, (np.asarray(coordinates)[:, 0].astype(np.int64), np.asarray(coordinates)[:, 1].astype(np.int64)), np.asarray(bulk, dtype=np.float64))
    if keep is not None:
        k = np.asarray(keep, dtype=np.float64)
        m *= np.outer(k, k)
    cov = np.sqrt(m.sum(-1))
    with np.errstate(divide="ignore", invalid="ignore"):
        m = m / cov[:, None] / cov[None, :]
    m[~np.isfinite(m)] = 0.0
    num, den = _window_sums(m, w)
    with np.errstate(divide="ignore", invalid="ignore"):
        score = num / den
    score[np.isnan(score)] = 1.0
    if discard_rows is not None and len(discard_rows):
        score[np.asarray(discard_rows, dtype=np.int64)] = 1.0
    idx = np.arange(n)
    ok = np.ones(n, dtype=bool)
    for d in range(1, order + 1):       # argrelextrema(np.less, order, mode="clip"): strict minima against +-d, edges clipped
        ok &= (score < score[np.clip(idx + d, 0, n - 1)]) & (score < score[np.clip(idx - d, 0, n - 1)])
    borders = np.flatnonzero(ok)
    return score, borders[(score[borders] < 0.999) & (score[borders] > 0)]


def gen_tad(eng, slot, chrom, cell_start, cell_end, window_ins, window_tad, res, keep=None, discard_rows=None, cells_per_call=128):
    """The return tuple of the reference's scTAD.gen_tad (scTAD.py:60-121): (chrom, sc_score, sc_border, sc_border_indice,
    bulk_score, bulk_tad_b) -- single-cell profiles from the device (gen_tad_scores), the pseudo-bulk profile on the host."""
    sc_score, sc_border, bulk = gen_tad_scores(eng, slot, cell_start, cell_end, window_ins, window_tad, res, keep, discard_rows, cells_per_call)
    n = sc_score.shape[1]
    bulk_score, bulk_tad_b = bulk_profile(eng.coordinates(slot), bulk, n, window_ins, window_tad, res, keep, discard_rows)
    sc_border_indice = np.empty(len(sc_border), dtype=object)
    for i, b in enumerate(sc_border):
        sc_border_indice[i] = np.flatnonzero(b)
    return chrom, sc_score.astype(np.float64), sc_border.astype(np.float64), sc_border_indice, bulk_score, bulk_tad_b
