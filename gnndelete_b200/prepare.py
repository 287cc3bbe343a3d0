"""Data preparation either side of the unlearning path: the edge split and the Df candidate
masks of the reference's ``prepare_dataset.py`` and their on-disk formats, on the device.

* ``d_{seed}.pkl`` - the reference pickles ``(dataset, data)`` with ``data`` a PyG ``Data`` holding the split
  (``prepare_dataset.py:190-198``).  Written here as ``(meta dict, GraphData)``; a file written by the reference
  unpickles (with PyG installed) to an object the trainers accept as is - they only use attribute access.
* ``df_{seed}.pt`` - ``torch.save({'out': bool[E], 'in': bool[E]})`` over the DIRECTED train edges (:262-265):
  ``in`` = edges inside the 2-hop enclosing subgraph of the test edges, ``out`` = the rest.  Same format here; the
  2-hop subgraph comes from the frontier-bitmap kernel (``masks.k_hop_subgraph``).
"""
from __future__ import annotations

import math
import os
import pickle

import torch

from . import masks as MK
from .data import GraphData
from .kg import negative_sampling_kg


def train_test_split_edges(data, val_ratio=0.05, test_ratio=0.1, two_hop_degree=None, kg=False, perm=None,
                           generator=None, permute_edge_type=False):
    """``train_test_split_edges_no_neg_adj_mask`` (``prepare_dataset.py:31-136``).

    ``data.edge_index`` (both directions for homogeneous graphs; head -> tail triples with ``data.edge_type`` for
    knowledge graphs) is split into a DIRECTED ``row < col`` train list plus test / val positives with as many
    sampled negatives.  ``perm`` supplies the edge permutation (the reference draws ``torch.randperm`` on the CPU
    generator; with ``two_hop_degree`` the edges whose 2-hop degree is below 50 come first, :52-61); the negatives
    are uniform random pairs (PyG's ``negative_sampling`` is not reproducible) or, for knowledge
    graphs, ``negative_sampling_kg``.

    Reference defect kept by default: on knowledge graphs the reference permutes ``row`` / ``col`` but slices the
    UN-permuted ``edge_type`` (:63-64 vs :78, :100, :118), so relation types no longer belong to their triples;
    ``permute_edge_type=True`` carries the types along with their edges instead."""
    row, col = data.edge_index
    dev = row.device
    edge_type = data.edge_type if kg else None
    out = data.clone()
    out.edge_index = None
    if not kg:                                                              # :45-50 upper triangular portion
        keep = row < col
        row, col = row[keep], col[keep]
    m = row.numel()
    n_v = int(math.floor(val_ratio * m))
    n_t = int(math.floor(test_ratio * m))
    if perm is None:
        if two_hop_degree is not None:                                      # :52-61
            low_mask = two_hop_degree.to(dev) < 50
            low, high = low_mask.nonzero().flatten(), (~low_mask).nonzero().flatten()
            low = low[torch.randperm(low.numel(), generator=generator, device=dev)]
            high = high[torch.randperm(high.numel(), generator=generator, device=dev)]
            perm = torch.cat([low, high])
        else:
            perm = torch.randperm(m, generator=generator, device=dev)
    perm = perm.to(dev)
    if perm.numel() != m:
        raise ValueError(f'perm has {perm.numel()} entries for {m} directed edges')
    row, col = row[perm], col[perm]
    if kg and permute_edge_type:
        edge_type = edge_type[perm]
    out.train_pos_edge_index = torch.stack([row[n_v + n_t:], col[n_v + n_t:]])          # :67, :86
    if kg:
        out.edge_index = out.train_pos_edge_index
        out.edge_type = out.train_edge_type = edge_type[n_v + n_t:]
    splits = (('test', 0, n_t), ('val', n_t, n_t + n_v))                     # :99-133
    for stage, lo, hi in splits:
        pos = torch.stack([row[lo:hi], col[lo:hi]])
        out[f'{stage}_pos_edge_index'] = pos
        if kg:
            out[f'{stage}_edge_type'] = edge_type[lo:hi]
            out[f'{stage}_neg_edge_index'] = negative_sampling_kg(pos, edge_type[lo:hi], generator)
        else:
            out[f'{stage}_neg_edge_index'] = torch.randint(0, int(data.num_nodes), (2, hi - lo), generator=generator, device=dev)
    return out


def df_candidate_masks(data, num_nodes=None):
    """``{'in': mask, 'out': ~mask}`` over the directed train edges (``prepare_dataset.py:203-215, 262-265``):
    ``mask`` = edges of the 2-hop subgraph around the endpoints of the test edges (``k_hop_subgraph`` with the
    source-to-target flow on the directed list)."""
    n = int(num_nodes if num_nodes is not None else data.num_nodes)
    seeds = data.test_pos_edge_index.flatten().unique()
    _, _, _, mask = MK.k_hop_subgraph(seeds, 2, data.train_pos_edge_index, num_nodes=n)
    return {'out': ~mask, 'in': mask}


def sample_df(candidates, df_size, num_edges, generator=None):
    """``delete_gnn.py:88-110``: ``df_size`` >= 100 is a count, otherwise a percentage of the directed train edges;
    a random subset of the candidate columns becomes ``df_mask`` (bool ``[E]``)."""
    size = int(df_size) if df_size >= 100 else int(df_size / 100 * num_edges)
    pool = candidates.nonzero().flatten()
    idx = pool[torch.randperm(pool.numel(), generator=generator, device=pool.device)[:size]]
    mask = torch.zeros(num_edges, dtype=torch.bool, device=candidates.device)
    mask[idx] = True
    return mask


def save_prepared(data_dir, dataset, seed, data, df_masks, meta=None):
    """Write ``d_{seed}.pkl`` and ``df_{seed}.pt`` under ``data_dir/dataset`` (tensors moved to the CPU)."""
    root = os.path.join(data_dir, dataset)
    os.makedirs(root, exist_ok=True)
    with open(os.path.join(root, f'd_{seed}.pkl'), 'wb') as f:
        pickle.dump((dict(meta or {}, name=dataset), data.clone().cpu()), f)
    torch.save({k: v.cpu() for k, v in df_masks.items()}, os.path.join(root, f'df_{seed}.pt'))


def load_prepared(data_dir, dataset, seed, df=None):
    """``delete_gnn.py:76-80, 97``: ``(dataset_meta, data)`` from ``d_{seed}.pkl`` and, with ``df`` = 'in' | 'out', the
    candidate mask from ``df_{seed}.pt``."""
    root = os.path.join(data_dir, dataset)
    with open(os.path.join(root, f'd_{seed}.pkl'), 'rb') as f:
        meta, data = pickle.load(f)
    if df is None:
        return meta, data
    masks = torch.load(os.path.join(root, f'df_{seed}.pt'))
    if df not in masks:
        raise KeyError(f"--df must be one of {sorted(masks)} (delete_gnn.py:86), got {df!r}")
    return meta, data, masks[df]
