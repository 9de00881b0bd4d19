"""Output sink with the reference's on-disk dataset layout (Impute.py:227-230, 277): one container per
chromosome `{temp_dir}/{chrom}_{name}.hdf5` holding `coordinates` int64 [P_c,2] and one float32 [P_c] dataset
`cell_{id}` per cell (gzip level 6).  `h5py` is used when importable; this image has neither h5py nor libhdf5, so
the fallback writes the same names / dtypes / shapes into `{chrom}_{name}.npz`."""
import os

import numpy as np

try:                                    # pragma: no cover - not installed in this image
    import h5py
except Exception:                       # noqa: BLE001
    h5py = None


class ChromSink:
    def __init__(self, temp_dir, chrom, name):
        self.base = os.path.join(temp_dir, "%s_%s" % (chrom, name))
        self.h5 = h5py.File(self.base + ".hdf5", "w") if h5py is not None else None
        self.data = {}

    def create_dataset(self, key, data, compress=False):
        if self.h5 is not None:
            kw = dict(compression="gzip", compression_opts=6) if compress else {}
            self.h5.create_dataset(key, data=data, **kw)
        else:
            self.data[key] = np.asarray(data)

    def close(self):
        if self.h5 is not None:
            self.h5.close()
        else:
            np.savez(self.base + ".npz", **self.data)


def read_container(temp_dir, chrom, name):
    base = os.path.join(temp_dir, "%s_%s" % (chrom, name))
    if h5py is not None and os.path.exists(base + ".hdf5"):
        with h5py.File(base + ".hdf5", "r") as f:
            return {k: np.asarray(f[k]) for k in f.keys()}
    with np.load(base + ".npz") as f:
        return {k: f[k] for k in f.files}


def read_num(temp_dir):
    """`num` of node_feats.hdf5 (Impute.py:150-152); `.npz` fallback with the same key."""
    p = os.path.join(temp_dir, "node_feats.hdf5")
    if h5py is not None and os.path.exists(p):
        with h5py.File(p, "r") as f:
            return np.asarray(f["num"])
    with np.load(os.path.join(temp_dir, "node_feats.npz")) as f:
        return np.asarray(f["num"])


def read_node_feats(temp_dir):
    """All datasets of `node_feats.hdf5` (Process.py:627,773,781-858) as a flat dict; groups are flattened with '/'
    (e.g. 'cell/0').  `.npz` twin with the same keys when h5py is unavailable."""
    p = os.path.join(temp_dir, "node_feats.hdf5")
    if h5py is not None and os.path.exists(p):
        out = {}

        def visit(name, obj):
            if hasattr(obj, "shape"):
                out[name] = np.asarray(obj)
        with h5py.File(p, "r") as f:
            f.visititems(visit)
        return out
    with np.load(os.path.join(temp_dir, "node_feats.npz"), allow_pickle=True) as f:
        return {k: f[k] for k in f.files}
