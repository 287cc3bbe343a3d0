"""Knowledge-graph helpers of the reference (``framework/utils.py:46-58``) on the device."""
from __future__ import annotations

import torch


@torch.no_grad()
def negative_sampling_kg(edge_index, edge_type, generator=None):
    """Per relation, permute the heads among that relation's edges (tails and relation
    types stay in place).  The reference loops over ``edge_type.unique()`` with a CPU
    ``randperm`` per relation; here one keyed sort permutes all relations at once.  Like
    the reference's, the result is random — parity runs supply the negatives instead."""
    dev = edge_index.device
    n = edge_type.numel()
    if n == 0:
        return edge_index.clone()
    noise = torch.rand(n, device=dev, generator=generator, dtype=torch.float64)
    shuffled = torch.argsort(edge_type.double() + noise * 0.999)     # random order inside each relation
    grouped = torch.argsort(edge_type, stable=True)                  # original positions, relation-major
    out = edge_index.clone()
    out[0, grouped] = edge_index[0, shuffled]
    return out
