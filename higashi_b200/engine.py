"""Host-side driver of the B200 scorer: feeds a reference-format model + data into libhsg.

`ScorerEngine` is what the mirrored reference classes (Higashi_backend.Modules) and
`Impute.impute_process` sit on.  It only marshals: numpy / torch buffers in, raw pointers to the C ABI.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ACT_NONE, ACT_SIGMOID, ACT_SOFTPLUS, PREC_BF16, PREC_FP32, PREC_TC16, PREC_TC16_HI  # noqa: F401


DEFAULT_PRECISION = PREC_TC16


def _np(t):
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t)


class ScorerEngine:
    def __init__(self, device: int = 0):
        self.ctx = _lib.Context(device)
        self.device = device
        self.P = 0
        self.sel_P: List[int] = []
        self.chroms: List[int] = []

    def close(self):
        self.ctx.close()

    # ------------------------------------------------------------------ model
    def load_model(self, sd: Dict[str, object], num: Sequence[int], feats: Optional[list] = None,
                   tables: Optional[list] = None, cell_feats=None,
                   embed_prefix: str = "encode1.static_nn"):
        """sd: reference state_dict (names of Modules.py); num = [n_cells, n_1, ..., n_C].

        feats[i] are the raw node features of branch i (cells, then chromosomes); tables[i], when given,
        is the already off-hooked [n_i, d] table of that branch (SparseEmbedding(embed), Modules.py:627)."""
        num = [int(v) for v in num]
        d = int(_np(sd["layer_norm1.weight"]).shape[0])
        cf = 0 if cell_feats is None else int(_np(cell_feats).shape[-1])
        self.num = num
        self.d = d
        self.ctx.set_dims(d, num[0], num[1:], cf)
        for key, slot in _lib.STATE_KEYS.items():
            if key in sd:
                self.ctx.set_weight(slot, _np(sd[key]).astype(np.float32))
        if cell_feats is not None:
            self.ctx.set_weight("W_CELL_FEATS", _np(cell_feats).astype(np.float32))
        self.tables = []
        for i in range(len(num)):
            if tables is not None and tables[i] is not None:
                t = _np(tables[i]).astype(np.float32)
                self.ctx.set_table(i, t)
            else:
                w = _np(sd["%s.wstack.%d.weight_list.0" % (embed_prefix, i)])
                b = _np(sd["%s.wstack.%d.bias_list.0" % (embed_prefix, i)])
                t = self.ctx.embed_table(i, _np(feats[i]), w, b, want_copy=True)
            self.tables.append(t)
        return self

    # ------------------------------------------------------------------ data
    def set_contacts(self, csr):
        self.ctx.set_contacts(csr.row_ptr, csr.col, csr.val)

    def set_neighbors(self, nbr_1based=None, w=None):
        """cell_neighbor_list / weights as the reference stores them: 1-based ids, self first."""
        if nbr_1based is None:
            self.ctx.set_neighbors(None, None)
        else:
            self.ctx.set_neighbors(np.asarray(nbr_1based, dtype=np.int64) - 1, np.asarray(w, dtype=np.float32))

    def set_pairs(self, chroms, min_bin, max_bin, skip=None):
        self.chroms = [int(c) for c in chroms]
        self.P, self.sel_P = self.ctx.set_pairs(self.chroms, min_bin, -1 if max_bin is None else max_bin, skip)
        return self.P

    def coordinates(self, slot):
        return self.ctx.coordinates(slot)

    def prepare(self, precision=PREC_FP32, local_transfer_range=0, activation=ACT_SIGMOID):
        self.ctx.prepare(precision, local_transfer_range, activation)

    # ------------------------------------------------------------------ compute
    def impute_to_host(self, cell_start, cell_end, out: Optional[np.ndarray] = None, half: bool = False) -> np.ndarray:
        """half=True: fp16 score block (hsg_impute_cells_host_f16), half the D2H bytes; widen with .astype(np.float32)."""
        n = cell_end - cell_start
        if out is None:
            out = np.empty((n, self.P), dtype=np.float16 if half else np.float32)
        if half:
            self.ctx.impute_cells_host_f16(cell_start, cell_end, out)
        else:
            self.ctx.impute_cells_host(cell_start, cell_end, out)
        return out

    def scorer_variant(self) -> int:
        """PREC_FP32 / PREC_TC16 / PREC_TC16_HI: the scorer hsg_prepare selected (after the fp16 range guard)."""
        return self.ctx.scorer_variant()

    def impute_to_device(self, cell_start, cell_end, out_tensor, stream_ptr: int = 0):
        """out_tensor: CUDA float32 torch tensor [n, P] (contiguous)."""
        assert out_tensor.is_cuda and out_tensor.is_contiguous() and out_tensor.dtype.is_floating_point
        assert out_tensor.numel() >= (cell_end - cell_start) * self.P
        if out_tensor.element_size() == 2:
            self.ctx.impute_cells_f16(cell_start, cell_end, out_tensor.data_ptr(), stream_ptr)
        else:
            self.ctx.impute_cells(cell_start, cell_end, out_tensor.data_ptr(), stream_ptr)
        return out_tensor

    def fix_cell(self, cell_1based, stream_ptr: int = 0):
        self.ctx.fix_cell(cell_1based, stream_ptr)

    def score_triplets(self, x_tensor, out_tensor, stream_ptr: int = 0):
        """x_tensor CUDA int64 [B,3] node ids; out_tensor CUDA float32 [B]."""
        assert x_tensor.is_cuda and x_tensor.is_contiguous() and out_tensor.is_cuda
        self.ctx.score_triplets(x_tensor.data_ptr(), x_tensor.shape[0], out_tensor.data_ptr(), stream_ptr)
        return out_tensor

    def insulation(self, slot, scores_tensor, window_bins, tad_order, keep=None, discard_rows=None, stream_ptr: int = 0):
        """Single-cell insulation scores + TAD boundary flags of one chromosome from the score block on the device
        (scTAD.py:91-104 without the n x n maps).  scores_tensor: CUDA float32 [n_batch, P]; discard_rows: bin indices."""
        assert scores_tensor.is_cuda and scores_tensor.is_contiguous() and scores_tensor.shape[1] == self.P
        n = int(self.num[1 + self.chroms[slot]])
        disc = None
        if discard_rows is not None:
            disc = np.zeros(n, dtype=np.uint8)
            disc[np.asarray(discard_rows, dtype=np.int64)] = 1
        return self.ctx.insulation(slot, scores_tensor.data_ptr(), scores_tensor.shape[0], n, window_bins, tad_order, keep, disc,
                                   stream_ptr)
