"""The shims live in oracle/ref_shims.py (shared with bench.py's reference arm); this name is kept for make_golden*.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_shims import *  # noqa: F401,F403,E402
from oracle.ref_shims import H5_STORE, REFERENCE_ROOT, import_reference, install  # noqa: F401,E402
