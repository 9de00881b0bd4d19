"""higashi_b200 -- B200-native implementation of Higashi's Hyper-SAGNN triplet-scoring hot path.

`libhsg.so` (C ABI, include/hsg.h) does the work; this package is the Python host side that mirrors the
reference's module API (Higashi_backend.Modules, Impute.impute_process)."""
import sys

__version__ = "0.1.0"


def install_aliases():
    """Make reference-style module paths resolve to the mirror, so that whole-model pickles written by the
    reference (`torch.save(model)`, class paths `Higashi_backend.Modules.*`) load here."""
    from . import Higashi_backend
    from .Higashi_backend import Functions, Modules, utils
    for prefix in ("Higashi_backend", "higashi.Higashi_backend"):
        sys.modules.setdefault(prefix, Higashi_backend)
        sys.modules.setdefault(prefix + ".Modules", Modules)
        sys.modules.setdefault(prefix + ".Functions", Functions)
        sys.modules.setdefault(prefix + ".utils", utils)
