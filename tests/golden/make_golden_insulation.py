"""Golden vectors for the single-cell insulation score / TAD boundary path.

The reference functions live in modules whose imports fail in this container (seaborn / matplotlib / h5py at import time), so the
UNMODIFIED source text of exactly these functions is extracted with `ast` from /root/reference and executed with numpy / scipy:

    Higashi_analysis/Higashi_TAD.py       insulation_score (:7-19), call_tads (:22-27)
    Higashi_analysis/Higashi_analysis.py  sqrt_norm (:581-588)

and the per-cell driver lines of scTAD.py:91-102 (matrix from coordinates + scores, mask, sqrt_norm, insulation_score,
score[discard_rows] = 1, call_tads) are replayed around them.  Run by hand:  python tests/golden/make_golden_insulation.py"""
import ast
import os

import numpy as np
from scipy.signal import argrelextrema  # noqa: F401  (used by the extracted call_tads)

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/higashi/Higashi_analysis"


def extract(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "argrelextrema": argrelextrema}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


def main():
    insulation_score, call_tads = extract(os.path.join(REF, "Higashi_TAD.py"), ["insulation_score", "call_tads"])
    (sqrt_norm,) = extract(os.path.join(REF, "Higashi_analysis.py"), ["sqrt_norm"])
    out = {}
    rng = np.random.Generator(np.random.PCG64(77))
    for tag, (n, band, w_ins, w_tad, n_cells) in {"a": (120, 40, 10, 6, 5), "b": (64, 64, 5, 3, 4), "c": (200, 25, 12, 10, 3)}.items():
        xs, ys = np.where(np.triu(np.ones((n, n)), 1) - np.triu(np.ones((n, n)), band + 1) > 0)      # 1 <= j - i <= band
        coords = np.stack([xs, ys], axis=1).astype(np.int64)
        keep = np.ones(n, dtype=np.uint8)
        keep[rng.choice(n, size=max(1, n // 15), replace=False)] = 0                                   # masked rows / columns
        mask = np.outer(keep, keep).astype(np.float64)
        discard = np.sort(rng.choice(n, size=max(1, n // 20), replace=False)).astype(np.int64)
        res = 1
        scores, borders = [], []
        m1 = np.zeros((n, n))
        all_proba = []
        for c in range(n_cells):
            # block-structured contact probabilities so that the insulation profile has real minima
            blk = rng.integers(8, 25)
            proba = (0.05 + 0.9 * ((xs // blk) == (ys // blk)) * rng.uniform(0.5, 1.0, size=len(xs)) +
                     0.05 * rng.random(len(xs))).astype(np.float32)
            if c == n_cells - 1:
                proba[:] = 0.0                                                                         # an empty cell: every score is nan -> 1
            all_proba.append(proba)
            m1 *= 0.0                                                     # scTAD.py:92-98
            m1[xs.astype("int"), ys.astype("int")] += proba
            m1 = m1 + m1.T
            temp = m1
            temp *= mask
            temp = sqrt_norm(temp)
            score = insulation_score(temp, windowsize=w_ins, res=res)      # scTAD.py:102
            score[discard] = 1.0
            border = call_tads(score, windowsize=w_tad * 2, res=res)       # windowsize_bin = int(windowsize / res / 2)
            scores.append(score)
            b = np.zeros(n, dtype=np.uint8); b[border] = 1
            borders.append(b)
        out["%s/coords" % tag] = coords
        out["%s/proba" % tag] = np.stack(all_proba)
        out["%s/keep" % tag] = keep
        out["%s/discard" % tag] = discard
        out["%s/cfg" % tag] = np.array([n, w_ins, w_tad], dtype=np.int64)
        out["%s/score" % tag] = np.stack(scores)
        out["%s/border" % tag] = np.stack(borders)
    np.savez_compressed(os.path.join(HERE, "insulation.npz"), **out)
    print("wrote insulation.npz", {k: v.shape for k, v in out.items() if k.startswith("a/")})


if __name__ == "__main__":
    main()
