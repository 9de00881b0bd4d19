"""Mirror of the hot-path helpers of higashi/Higashi_backend/utils.py."""
import json

import numpy as np
import torch


def get_config(config_path="./config.jSON"):
    """utils.py:25"""
    with open(config_path, "r") as c:
        return json.load(c)


def np2tensor_hyper(vec, dtype):
    """utils.py:57"""
    if len(vec.shape) == 1:
        return [torch.as_tensor(v, dtype=dtype) for v in vec]
    return torch.as_tensor(vec, dtype=dtype)


def skip_start_end(config, chrom="chr1"):
    """utils.py:178 / Impute.py:29: bins within +-100 kb of the `acen` cytobands of `chrom`."""
    res = config["resolution"]
    start, end = [], []
    if "cytoband_path" in config:
        with open(config["cytoband_path"]) as f:
            for line in f:
                if line.startswith("#") or not line.strip():
                    continue
                t = line.rstrip("\n").split("\t")
                if len(t) >= 5 and t[0] == chrom and t[4] == "acen":
                    start.append(int(np.floor((int(t[1]) - 100000) / res)))
                    end.append(int(np.ceil((int(t[2]) + 100000) / res)))
    return np.asarray(start, dtype="int"), np.asarray(end, dtype="int")


def generate_binpair(start, end, min_bin_, max_bin_, not_use_set=None):
    """utils.py:195 / Impute.py:49 -- enumerated on the device (K3, bit-exact), returned as int64 [P,2] of 1-based
    global node ids like the reference."""
    from .. import _lib
    n = int(end - start)
    ctx = _lib.Context(torch.cuda.current_device() if torch.cuda.is_available() else 0)
    try:
        ctx.set_dims(64, max(1, int(start)), [max(1, n)], 1)
        skip = [[0, int(b) - int(start), int(b) - int(start) + 1] for b in sorted(not_use_set or [])]
        ctx.set_pairs([0], int(min_bin_), int(min(max_bin_, n + 1)), skip)
        coords = ctx.coordinates(0)
    finally:
        ctx.close()
    return coords + int(start) + 1
