"""Mirror of higashi/Higashi_backend/Functions.py (the part the scorer path uses)."""
import torch


def swish(x):
    """Functions.py:14 -- the model-wide activation (Modules.py:27).  Host-side helper only: inside the scorer
    the activation is fused into the CUDA kernels."""
    return x * torch.sigmoid(x)


def get_non_pad_mask(seq):
    assert seq.dim() == 2
    return seq.ne(0).type(torch.float).unsqueeze(-1)
