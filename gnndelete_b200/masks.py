"""Deletion-mask pipeline on the device.

``k_hop_subgraph`` / ``to_undirected`` keep the call shapes of the PyG utilities the
reference uses (``delete_gnn.py:128-151, 175-182``); ``build_unlearning_data`` is the
mask-building preamble of ``delete_gnn.py:113-189`` run entirely with the CUDA kernels
in ``csrc/masks.cu``.  Results are bit-exact against the oracle.
"""
from __future__ import annotations

import torch

from . import _lib as L


def khop_masks(edge_index, seed_edge_mask, num_hops, num_nodes):
    """(edge_mask bool [E], node_mask bool [N]) of the ``num_hops`` subgraph seeded by the
    endpoints of ``edge_index[:, seed_edge_mask]`` (flow='source_to_target')."""
    dev = edge_index.device
    E, N = edge_index.shape[1], int(num_nodes)
    src, dst = edge_index[0].contiguous(), edge_index[1].contiguous()
    sel = seed_edge_mask.to(torch.uint8).contiguous()
    edge_mask = torch.zeros(max(E, 1), dtype=torch.uint8, device=dev)
    node_mask = torch.zeros(max(N, 1), dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = L.load().gd_khop_workspace_bytes(N)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    L.call('gd_khop_masks', L.ptr(src, 'i64'), L.ptr(dst, 'i64'), E, N, L.ptr(sel), int(num_hops),
           L.ptr(edge_mask), L.ptr(node_mask), L.ptr(status), L.ptr(ws), nbytes, L.stream())
    bad = int(status.item())
    if bad:
        raise ValueError(f'edge_index holds {bad} entries with endpoints outside [0, {N})')
    return edge_mask[:E].bool(), node_mask[:N].bool()


def k_hop_subgraph(node_idx, num_hops, edge_index, num_nodes=None):
    """PyG-shaped wrapper: ``(subset, edge_index[:, edge_mask], inv, edge_mask)`` for
    ``relabel_nodes=False, flow='source_to_target'``.  ``subset`` is returned as the
    endpoints-of-induced-edges node set plus the seeds themselves."""
    n = int(num_nodes) if num_nodes is not None else int(edge_index.max()) + 1
    dev = edge_index.device
    node_idx = torch.as_tensor(node_idx, device=dev).flatten().long()
    # express the seed node set as a seed-edge mask over an augmented list of self loops
    loops = torch.stack([node_idx, node_idx])
    aug = torch.cat([edge_index, loops], 1)
    sel = torch.zeros(aug.shape[1], dtype=torch.bool, device=dev)
    sel[edge_index.shape[1]:] = True
    edge_mask, node_mask = khop_masks(aug, sel, num_hops, n)
    edge_mask = edge_mask[:edge_index.shape[1]]
    subset = node_mask.nonzero().squeeze(1)      # seeds carry self loops, so they are always included
    inv = torch.searchsorted(subset, node_idx)
    return subset, edge_index[:, edge_mask], inv, edge_mask


def to_undirected(edge_index, edge_attrs=None, num_nodes=None):
    """``to_undirected(edge_index, [a, b])`` with int attributes summed over duplicates."""
    dev = edge_index.device
    E = edge_index.shape[1]
    n = int(num_nodes) if num_nodes is not None else (int(edge_index.max()) + 1 if E else 0)
    attrs = list(edge_attrs or [])
    if len(attrs) > 2:
        raise NotImplementedError('at most two edge attributes (delete_gnn.py:175 passes two)')
    a = [t.to(torch.int32).contiguous() for t in attrs] + [None, None]
    src, dst = edge_index[0].contiguous(), edge_index[1].contiguous()
    out_row = torch.empty(max(2 * E, 1), dtype=torch.int64, device=dev)
    out_col = torch.empty(max(2 * E, 1), dtype=torch.int64, device=dev)
    outs = [torch.empty(max(2 * E, 1), dtype=torch.int32, device=dev) if t is not None else None for t in a[:2]]
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    nbytes = L.load().gd_to_undirected_workspace_bytes(E)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    L.call('gd_to_undirected', L.ptr(src, 'i64'), L.ptr(dst, 'i64'), E, n, L.ptr(a[0]), L.ptr(a[1]),
           L.ptr(out_row), L.ptr(out_col), L.ptr(outs[0]), L.ptr(outs[1]), L.ptr(count), L.ptr(ws), nbytes,
           L.stream())
    m = int(count.item())
    sym = torch.stack([out_row[:m], out_col[:m]])
    if edge_attrs is None:
        return sym
    return sym, [o[:m] for o in outs[:len(attrs)]]


def build_unlearning_data(data, df_mask, num_edge_type=None):
    """``delete_gnn.py:113-189`` on the device: ``directed_df_edge_index``, 2-hop edge
    mask ``sdf_mask``, 1-/2-hop node masks (on the DIRECTED list — PyG quirk, SURVEY.md
    §9.5), then symmetrisation carrying ``df_mask`` / ``sdf_mask`` (homogeneous) or
    reverse-edge doubling with types ``+ num_edge_type`` (knowledge graphs)."""
    out = data.clone()
    ei = data.train_pos_edge_index
    n = data.num_nodes
    df_mask = df_mask.to(ei.device)
    out.directed_df_edge_index = ei[:, df_mask]
    if num_edge_type is not None:
        out.directed_df_edge_type = data.train_edge_type[df_mask]
    two_hop_mask, s2 = khop_masks(ei, df_mask, 2, n)
    _, s1 = khop_masks(ei, df_mask, 1, n)
    out.sdf_node_1hop_mask = s1
    out.sdf_node_2hop_mask = s2
    if num_edge_type is not None:
        out.edge_index = torch.cat([ei, ei.flip(0)], 1)
        out.edge_type = torch.cat([data.train_edge_type, data.train_edge_type + num_edge_type])
        two_hop_mask = two_hop_mask.repeat(2)
        df_sym = df_mask.repeat(2)
    else:
        sym, (df_i, sdf_i) = to_undirected(ei, [df_mask.int(), two_hop_mask.int()], num_nodes=n)
        two_hop_mask = sdf_i.bool()
        df_sym = df_i.bool()
        out.train_pos_edge_index = sym
        out.edge_index = sym
    out.sdf_mask = two_hop_mask
    out.df_mask = df_sym
    out.dr_mask = ~df_sym
    return out
