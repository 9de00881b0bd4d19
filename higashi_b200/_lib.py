"""ctypes binding of libhsg.so (C ABI in include/hsg.h).  No torch types cross this boundary.

The library is built in-tree by `__graft_entry__.build()` / `make -C higashi_b200/csrc` and must be
present: there is no CPU fallback, every failure raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HSG_LIB_PATH", os.path.join(_HERE, "libhsg.so"))

# enum hsg_weight (include/hsg.h) -- order matters
WEIGHT_SLOTS = [
    "W_QS", "W_KS", "W_VS", "W_FC1", "W_FC2",
    "W_ALN1_G", "W_ALN1_B", "W_ALN2_G", "W_ALN2_B", "W_ALN3_G", "W_ALN3_B",
    "W_PFF_LN_G", "W_PFF_LN_B", "W_PFF_W0", "W_PFF_B0", "W_PFF_W1", "W_PFF_B1",
    "W_LN1_G", "W_LN1_B", "W_LN2_G", "W_LN2_B",
    "W_CLS_W0", "W_CLS_B0", "W_CLS_W1", "W_CLS_B1",
    "W_VAR_W0", "W_VAR_B0", "W_VAR_W1", "W_VAR_B1",
    "W_PRB_W0", "W_PRB_B0", "W_PRB_W1", "W_PRB_B1",
    "W_SAGE_W", "W_SAGE_B",
    "W_XP_W0", "W_XP_B0", "W_XP_W1", "W_XP_B1",
    "W_XP2_W0", "W_XP2_B0", "W_XP2_W1", "W_XP2_B1",
    "W_XP3_W0", "W_XP3_B0", "W_XP3_W1", "W_XP3_B1",
    "W_ATTR", "W_CELL_FEATS",
]
SLOT = {n: i for i, n in enumerate(WEIGHT_SLOTS)}

# reference state_dict key -> slot  (Modules.py; see include/hsg.h)
STATE_KEYS = {
    "encode1.mul_head_attn.w_qs.weight": "W_QS",
    "encode1.mul_head_attn.w_ks.weight": "W_KS",
    "encode1.mul_head_attn.w_vs.weight": "W_VS",
    "encode1.mul_head_attn.fc1.w_stack.0.weight": "W_FC1",
    "encode1.mul_head_attn.fc2.w_stack.0.weight": "W_FC2",
    "encode1.mul_head_attn.layer_norm1.weight": "W_ALN1_G", "encode1.mul_head_attn.layer_norm1.bias": "W_ALN1_B",
    "encode1.mul_head_attn.layer_norm2.weight": "W_ALN2_G", "encode1.mul_head_attn.layer_norm2.bias": "W_ALN2_B",
    "encode1.mul_head_attn.layer_norm3.weight": "W_ALN3_G", "encode1.mul_head_attn.layer_norm3.bias": "W_ALN3_B",
    "encode1.pff_n1.layer_norm.weight": "W_PFF_LN_G", "encode1.pff_n1.layer_norm.bias": "W_PFF_LN_B",
    "encode1.pff_n1.w_stack.0.weight": "W_PFF_W0", "encode1.pff_n1.w_stack.0.bias": "W_PFF_B0",
    "encode1.pff_n1.w_stack.1.weight": "W_PFF_W1", "encode1.pff_n1.w_stack.1.bias": "W_PFF_B1",
    "layer_norm1.weight": "W_LN1_G", "layer_norm1.bias": "W_LN1_B",
    "layer_norm2.weight": "W_LN2_G", "layer_norm2.bias": "W_LN2_B",
    "pff_classifier.w_stack.0.weight": "W_CLS_W0", "pff_classifier.w_stack.0.bias": "W_CLS_B0",
    "pff_classifier.w_stack.1.weight": "W_CLS_W1", "pff_classifier.w_stack.1.bias": "W_CLS_B1",
    "pff_classifier_var.w_stack.0.weight": "W_VAR_W0", "pff_classifier_var.w_stack.0.bias": "W_VAR_B0",
    "pff_classifier_var.w_stack.1.weight": "W_VAR_W1", "pff_classifier_var.w_stack.1.bias": "W_VAR_B1",
    "pff_classifier_proba.w_stack.0.weight": "W_PRB_W0", "pff_classifier_proba.w_stack.0.bias": "W_PRB_B0",
    "pff_classifier_proba.w_stack.1.weight": "W_PRB_W1", "pff_classifier_proba.w_stack.1.bias": "W_PRB_B1",
    "encode1.dynamic_nn.nn.weight": "W_SAGE_W", "encode1.dynamic_nn.nn.bias": "W_SAGE_B",
    "extra_proba.w_stack.0.weight": "W_XP_W0", "extra_proba.w_stack.0.bias": "W_XP_B0",
    "extra_proba.w_stack.1.weight": "W_XP_W1", "extra_proba.w_stack.1.bias": "W_XP_B1",
    "extra_proba2.w_stack.0.weight": "W_XP2_W0", "extra_proba2.w_stack.0.bias": "W_XP2_B0",
    "extra_proba2.w_stack.1.weight": "W_XP2_W1", "extra_proba2.w_stack.1.bias": "W_XP2_B1",
    "extra_proba3.w_stack.0.weight": "W_XP3_W0", "extra_proba3.w_stack.0.bias": "W_XP3_B0",
    "extra_proba3.w_stack.1.weight": "W_XP3_W1", "extra_proba3.w_stack.1.bias": "W_XP3_B1",
    "attribute_dict_embedding.weight": "W_ATTR",
}

PREC_FP32, PREC_TC16 = 0, 1
PREC_BF16 = PREC_TC16   # historical alias
PREC_TC16_HI = 2
ACT_NONE, ACT_SIGMOID, ACT_SOFTPLUS = 0, 1, 2

# every symbol include/hsg.h declares (tests check that the library exports all of them)
EXPORTS = [
    "hsg_abi_version", "hsg_last_error", "hsg_create", "hsg_destroy", "hsg_set_dims", "hsg_set_weight",
    "hsg_embed_table", "hsg_set_table", "hsg_set_contacts", "hsg_stage_contacts", "hsg_commit_contacts", "hsg_set_neighbors", "hsg_build_neighbors",
    "hsg_set_pairs", "hsg_get_coordinates", "hsg_prepare", "hsg_impute_cells", "hsg_impute_cells_host",
    "hsg_fix_cell", "hsg_score_triplets", "hsg_launch_count", "hsg_last_timing", "hsg_set_profiling",
    "hsg_train_bind", "hsg_train_set_features", "hsg_score_fwd", "hsg_score_bwd", "hsg_train_options", "hsg_score_heads",
    "hsg_train_bind_recon", "hsg_recon_loss", "hsg_recon_apply", "hsg_train_set_dropout",
    "hsg_score_set_head_grads", "hsg_sample_batch", "hsg_sample_batch_begin", "hsg_sample_batch_end", "hsg_sampled_fetch",
    "hsg_insulation", "hsg_impute_cells_f16", "hsg_impute_cells_host_f16", "hsg_refresh_contacts", "hsg_scorer_variant", "hsg_sample_set_removal_draws",
    "hsg_debug_tile_table",
]


class HsgError(RuntimeError):
    pass


_lib = None


def load():
    """dlopen libhsg.so and declare the prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HsgError("libhsg.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `make -C higashi_b200/csrc`; there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    p, i32, i64, f32p = C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_float)
    i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    lib.hsg_abi_version.restype = i32
    lib.hsg_last_error.restype = C.c_char_p
    lib.hsg_create.argtypes = [C.POINTER(p), i32]
    lib.hsg_destroy.argtypes = [p]
    lib.hsg_set_dims.argtypes = [p, i32, i32, i32, i32p, i32]
    lib.hsg_set_weight.argtypes = [p, i32, f32p, i64]
    lib.hsg_embed_table.argtypes = [p, i32, f32p, i64, i32, f32p, f32p, f32p]
    lib.hsg_set_table.argtypes = [p, i32, f32p, i64]
    lib.hsg_set_contacts.argtypes = [p, i64p, i32p, f32p, i64, i64]
    lib.hsg_stage_contacts.argtypes = [p, i64p, i32p, f32p, i64, i64]
    lib.hsg_commit_contacts.argtypes = [p]
    lib.hsg_set_neighbors.argtypes = [p, i32p, f32p, i32]
    lib.hsg_build_neighbors.argtypes = [p, f32p, i32, i32, i32p, f32p]
    lib.hsg_set_pairs.argtypes = [p, i32, i32p, i32, i32, i32p, i32, i64p, i64p]
    lib.hsg_get_coordinates.argtypes = [p, i32, i64p]
    lib.hsg_prepare.argtypes = [p, i32, i32, i32]
    lib.hsg_impute_cells.argtypes = [p, i32, i32, p, p]
    lib.hsg_impute_cells_host.argtypes = [p, i32, i32, p]
    lib.hsg_impute_cells_f16.argtypes = [p, i32, i32, p, p]
    lib.hsg_impute_cells_host_f16.argtypes = [p, i32, i32, p]
    lib.hsg_refresh_contacts.argtypes = [p, i32, i32, p, p, i64]
    lib.hsg_scorer_variant.argtypes = [p, i32p]
    lib.hsg_sample_set_removal_draws.argtypes = [p, f32p, i32]
    lib.hsg_fix_cell.argtypes = [p, i32, p]
    lib.hsg_score_triplets.argtypes = [p, p, i64, p, p]
    lib.hsg_launch_count.argtypes = [p]
    lib.hsg_launch_count.restype = i64
    lib.hsg_last_timing.argtypes = [p, f32p]
    lib.hsg_set_profiling.argtypes = [p, i32]
    lib.hsg_train_bind.argtypes = [p, p, p, i64, i64p, i64p, i64p]
    lib.hsg_train_set_features.argtypes = [p, i32, f32p, i64, i32]
    lib.hsg_score_fwd.argtypes = [p, i32, i32, i64p, i32, i32p, i32p, f32p, i32, f32p, f32p, i32, i32, f32p, f32p, p]
    lib.hsg_score_set_head_grads.argtypes = [p, p, p, p, p]
    lib.hsg_sample_batch.argtypes = [p, i64p, f32p, i32, i32, i32, C.c_float, i32, i32, i32, C.c_float, C.c_uint64, i32p, i64p, p]
    lib.hsg_sample_batch_begin.argtypes = [p, i64p, f32p, i32, i32, i32, C.c_float, i32, i32, i32, C.c_float, C.c_uint64]
    lib.hsg_sample_batch_end.argtypes = [p, i32p, i64p, p]
    lib.hsg_sampled_fetch.argtypes = [p, i64p, f32p, f32p, i32p, i32p, i32p, f32p, p]
    lib.hsg_score_bwd.argtypes = [p, p]
    lib.hsg_train_options.argtypes = [p, f32p, C.c_float]
    lib.hsg_score_heads.argtypes = [p, f32p, f32p, f32p, f32p, p]
    lib.hsg_train_bind_recon.argtypes = [p, i64, i64]
    lib.hsg_recon_loss.argtypes = [p, f32p, f32p, p]
    lib.hsg_recon_apply.argtypes = [p, C.c_float, p]
    lib.hsg_train_set_dropout.argtypes = [p, i32, C.c_uint64]
    lib.hsg_insulation.argtypes = [p, i32, p, i32, i32, i32, p, p, f32p, p, p]
    lib.hsg_debug_tile_table.argtypes = [C.POINTER(C.c_int32), i64, i32, C.POINTER(C.c_int32), i64]
    for name in EXPORTS:
        if name not in ("hsg_last_error", "hsg_launch_count", "hsg_abi_version"):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int32))


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(C.POINTER(C.c_int64))


class Context:
    """Owns one hsg_ctx.  Thin: every method is one C call."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.h = C.c_void_p()
        self._check(self.lib.hsg_create(C.byref(self.h), int(device)))
        self.P = 0
        self.sel_P = []

    def _check(self, rc):
        if rc != 0:
            raise HsgError(self.lib.hsg_last_error().decode("utf-8", "replace"))

    def close(self):
        if self.h:
            self.lib.hsg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- model
    def set_dims(self, d_model, n_cells, bins, cf):
        b, bp = _i32(bins)
        self.d, self.n_cells, self.bins, self.cf = int(d_model), int(n_cells), [int(x) for x in bins], int(cf)
        self._check(self.lib.hsg_set_dims(self.h, int(d_model), int(n_cells), len(b), bp, int(cf)))

    def set_weight(self, slot, arr):
        a, ap = _f32(np.asarray(arr).reshape(-1))
        self._check(self.lib.hsg_set_weight(self.h, SLOT[slot] if isinstance(slot, str) else int(slot), ap, a.size))

    def embed_table(self, branch, x, w, b, want_copy=False):
        x, xp = _f32(x); w, wp = _f32(w); b, bp = _f32(b)
        out = np.empty((x.shape[0], w.shape[0]), np.float32) if want_copy else None
        op = out.ctypes.data_as(C.POINTER(C.c_float)) if want_copy else None
        self._check(self.lib.hsg_embed_table(self.h, int(branch), xp, x.shape[0], x.shape[1], wp, bp, op))
        return out

    def set_table(self, branch, table):
        t, tp = _f32(table)
        self._check(self.lib.hsg_set_table(self.h, int(branch), tp, t.shape[0]))

    # ---- data
    def set_contacts(self, row_ptr, col, val):
        r, rp = _i64(row_ptr); c, cp = _i32(col); v, vp = _f32(val)
        self._check(self.lib.hsg_set_contacts(self.h, rp, cp, vp, r.size - 1, c.size))

    def stage_contacts(self, row_ptr, col, val):
        """Asynchronous upload of the NEXT contact matrix (own stream, second buffer set); the arrays must stay alive and
        unchanged until commit_contacts()."""
        r, rp = _i64(row_ptr); c, cp = _i32(col); v, vp = _f32(val)
        self._staged_keep = (r, c, v)
        self._check(self.lib.hsg_stage_contacts(self.h, rp, cp, vp, r.size - 1, c.size))

    def commit_contacts(self):
        self._check(self.lib.hsg_commit_contacts(self.h))
        self._staged_keep = None

    def set_neighbors(self, nbr0, w):
        if nbr0 is None:
            self._check(self.lib.hsg_set_neighbors(self.h, None, None, 1))
            return
        n, np_ = _i32(nbr0); w, wp = _f32(w)
        self._check(self.lib.hsg_set_neighbors(self.h, np_, wp, n.shape[1]))

    def build_neighbors(self, embed, k1):
        e, ep = _f32(embed)
        nbr = np.empty((e.shape[0], k1), np.int32); w = np.empty((e.shape[0], k1), np.float32)
        self._check(self.lib.hsg_build_neighbors(self.h, ep, e.shape[1], int(k1), nbr.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 w.ctypes.data_as(C.POINTER(C.c_float))))
        return nbr, w

    def set_pairs(self, chroms, min_bin, max_bin, skip=None):
        ch, chp = _i32(chroms)
        if skip is None or len(skip) == 0:
            sk, skp, ns = None, None, 0
        else:
            sk, skp = _i32(np.asarray(skip).reshape(-1, 3)); ns = sk.shape[0]
        P = C.c_int64(0)
        per = np.zeros(len(ch), np.int64)
        self._check(self.lib.hsg_set_pairs(self.h, len(ch), chp, int(min_bin), int(max_bin), skp, ns, C.byref(P),
                                           per.ctypes.data_as(C.POINTER(C.c_int64))))
        self.P, self.sel_P = int(P.value), [int(v) for v in per]
        return self.P, self.sel_P

    def coordinates(self, sel_slot):
        out = np.empty((self.sel_P[sel_slot], 2), np.int64)
        self._check(self.lib.hsg_get_coordinates(self.h, int(sel_slot), out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    # ---- compute
    def prepare(self, precision=PREC_FP32, local_transfer_range=0, activation=ACT_SIGMOID):
        self._check(self.lib.hsg_prepare(self.h, int(precision), int(local_transfer_range), int(activation)))

    def impute_cells(self, cell_start, cell_end, out_dev_ptr, stream=0):
        self._check(self.lib.hsg_impute_cells(self.h, int(cell_start), int(cell_end), C.c_void_p(out_dev_ptr), C.c_void_p(stream)))

    def impute_cells_host(self, cell_start, cell_end, out_host: np.ndarray):
        assert out_host.dtype == np.float32 and out_host.flags["C_CONTIGUOUS"]
        self._check(self.lib.hsg_impute_cells_host(self.h, int(cell_start), int(cell_end), C.c_void_p(out_host.ctypes.data)))

    def impute_cells_host_ptr(self, cell_start, cell_end, host_ptr, half=False):
        f = self.lib.hsg_impute_cells_host_f16 if half else self.lib.hsg_impute_cells_host
        self._check(f(self.h, int(cell_start), int(cell_end), C.c_void_p(host_ptr)))

    def impute_cells_f16(self, cell_start, cell_end, out_dev_ptr, stream=0):
        self._check(self.lib.hsg_impute_cells_f16(self.h, int(cell_start), int(cell_end), C.c_void_p(out_dev_ptr), C.c_void_p(stream)))

    def impute_cells_host_f16(self, cell_start, cell_end, out_host: np.ndarray):
        assert out_host.dtype == np.float16 and out_host.flags["C_CONTIGUOUS"]
        self._check(self.lib.hsg_impute_cells_host_f16(self.h, int(cell_start), int(cell_end), C.c_void_p(out_host.ctypes.data)))

    def refresh_contacts_ptr(self, cell_start, cell_end, col_ptr, val_ptr, nnz_range):
        """Asynchronous re-upload of the stored entries of the rows of cells [cell_start, cell_end) (host pointers)."""
        self._check(self.lib.hsg_refresh_contacts(self.h, int(cell_start), int(cell_end), C.c_void_p(col_ptr), C.c_void_p(val_ptr), int(nnz_range)))

    def scorer_variant(self):
        v = C.c_int32(0)
        self._check(self.lib.hsg_scorer_variant(self.h, C.byref(v)))
        return int(v.value)

    def fix_cell(self, cell_1based, stream=0):
        self._check(self.lib.hsg_fix_cell(self.h, int(cell_1based), C.c_void_p(stream)))

    def score_triplets(self, x_dev_ptr, B, out_dev_ptr, stream=0):
        self._check(self.lib.hsg_score_triplets(self.h, C.c_void_p(x_dev_ptr), int(B), C.c_void_p(out_dev_ptr), C.c_void_p(stream)))

    def insulation(self, sel_slot, scores_dev_ptr, n_batch, n, window_bins, tad_order, keep=None, discard=None, stream=0):
        """scores on the DEVICE ([n_batch, P_total] block of impute_cells) -> (score float32 [n_batch, n], border uint8 [n_batch, n])."""
        score = np.empty((n_batch, n), dtype=np.float32)
        border = np.empty((n_batch, n), dtype=np.uint8)
        kp = dp = None
        if keep is not None:
            keep = np.ascontiguousarray(keep, dtype=np.uint8)
            assert keep.shape == (n,)
            kp = keep.ctypes.data_as(C.c_void_p)
        if discard is not None:
            discard = np.ascontiguousarray(discard, dtype=np.uint8)
            assert discard.shape == (n,)
            dp = discard.ctypes.data_as(C.c_void_p)
        self._check(self.lib.hsg_insulation(self.h, int(sel_slot), C.c_void_p(scores_dev_ptr), int(n_batch), int(window_bins),
                                            int(tad_order), kp, dp, score.ctypes.data_as(C.POINTER(C.c_float)),
                                            border.ctypes.data_as(C.c_void_p), C.c_void_p(stream)))
        return score, border

    # ---- introspection
    def launch_count(self):
        return int(self.lib.hsg_launch_count(self.h))

    def set_profiling(self, on):
        self._check(self.lib.hsg_set_profiling(self.h, 1 if on else 0))

    def last_timing(self):
        ms = (C.c_float * 4)()
        self._check(self.lib.hsg_last_timing(self.h, ms))
        return [float(v) for v in ms]
