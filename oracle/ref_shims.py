"""Import shims that let the UNMODIFIED reference run without its optional dependencies (test infrastructure).

Used by `tests/golden/make_golden.py` (run by hand in the container that has /root/reference) and by the reference arm of
`bench.py` (`--impl reference`), which runs the reference's own modules from `baseline/_ref` -- a `pip install --target` of the
unmodified reference, git-ignored, shipped to the GPU box next to the built library -- and falls back to the oracle port when that
install is absent.  Nothing in `higashi_b200/`, the `-m gpu` tests or `smoke()` imports this file.

Shims: a stub `cooler`, a dict-backed fake `h5py`, and the
`scipy.sparse.csr.get_csr_submatrix` alias that newer scipy dropped.
"""
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference_root():
    """Directory that holds the reference's `higashi/` package: $HIGASHI_REFERENCE, the checkout of the build container, or the
    pip --target install under baseline/_ref (the only one that exists on the GPU box)."""
    cands = [os.environ.get("HIGASHI_REFERENCE"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "higashi", "Impute.py")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_reference_root()

# ---------------------------------------------------------------- fake h5py
H5_STORE = {}          # path -> {dataset name -> ndarray}


class _FakeGroup(dict):
    def create_dataset(self, name, data=None, **kw):
        self[name] = np.array(data)
        return self[name]

    def create_group(self, name):
        self[name] = _FakeGroup()
        return self[name]


class _FakeFile(_FakeGroup):
    def __init__(self, path, mode="r"):
        super().__init__()
        self.path = path
        if mode in ("w",):
            H5_STORE[path] = self
        else:
            if path not in H5_STORE:
                raise OSError("fake h5py: no such file %s" % path)
            self.update(H5_STORE[path])
            H5_STORE[path] = self

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def install():
    if "cooler" not in sys.modules:
        sys.modules["cooler"] = types.ModuleType("cooler")
    if "h5py" not in sys.modules:
        m = types.ModuleType("h5py")
        m.File = _FakeFile
        sys.modules["h5py"] = m
    import scipy.sparse
    import scipy.sparse._sparsetools as st
    try:
        import scipy.sparse.csr as csr_mod
    except Exception:                                    # pragma: no cover
        csr_mod = types.ModuleType("scipy.sparse.csr")
        sys.modules["scipy.sparse.csr"] = csr_mod
    if not hasattr(csr_mod, "get_csr_submatrix"):
        csr_mod.get_csr_submatrix = st.get_csr_submatrix
    p = os.path.join(REFERENCE_ROOT, "higashi")
    if p not in sys.path:
        sys.path.insert(0, p)


_REF_CACHE = None


def import_reference():
    """Returns (Modules, Impute, Higashi_wrapper) of the reference.

    `higashi_b200.install_aliases()` may have registered the host mirror under the reference's module names
    (`Higashi_backend.Modules`, ...) in this process; those entries are set aside while the reference's own files are imported
    and put back afterwards, so both can live in one interpreter (the reference modules stay reachable through the returned
    objects only)."""
    global _REF_CACHE
    if _REF_CACHE is not None:
        return _REF_CACHE
    install()
    import warnings
    warnings.filterwarnings("ignore")
    names = [n for n in list(sys.modules) if n == "Higashi_backend" or n.startswith("Higashi_backend.") or n in ("Impute", "Higashi_wrapper")]
    saved = {n: sys.modules.pop(n) for n in names}
    try:
        import Higashi_backend.Modules as RM
        import Impute as RI
        import Higashi_wrapper as RW
    finally:
        for n in [n for n in list(sys.modules) if n == "Higashi_backend" or n.startswith("Higashi_backend.") or n in ("Impute", "Higashi_wrapper")]:
            if n in saved:
                sys.modules[n] = saved[n]
            elif saved:
                del sys.modules[n]
    root = os.path.realpath(REFERENCE_ROOT)
    assert os.path.realpath(RM.__file__).startswith(root), "reference import resolved to %s" % RM.__file__
    _REF_CACHE = (RM, RI, RW)
    return _REF_CACHE
