"""TEST INFRASTRUCTURE — CPU restatement of the PyTorch-Geometric operators the
reference's hot path calls.  Never imported by the product package.

PARITY UNPINNED: PyG (unpinned by the reference; API usage brackets it to
2.0.3 … 2.2.x) is not installed here and its source is not under
``/root/reference``; the reference holds no golden vectors.  Each function
restates PyG's *published* default-path algorithm for the constructor arguments
the reference uses, as an eager op sequence (index_select gather -> message ->
index_add_ scatter) so that it doubles as the CPU timing baseline.  Call sites
that fix the arguments are cited per function (paths relative to /root/reference).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- utils
def maybe_num_nodes(edge_index, num_nodes=None):
    if num_nodes is not None:
        return int(num_nodes)
    return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0


def add_remaining_self_loops(edge_index, num_nodes):
    """GCNConv / GATConv default ``add_self_loops=True``: existing (v,v) entries
    are dropped and one (v,v) per node is appended AT THE END with weight 1."""
    keep = edge_index[0] != edge_index[1]
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index[:, keep], loop.unsqueeze(0).repeat(2, 1)], dim=1)


def is_undirected(edge_index, num_nodes=None):
    n = maybe_num_nodes(edge_index, num_nodes)
    a = torch.unique(edge_index[0] * n + edge_index[1])
    b = torch.unique(edge_index[1] * n + edge_index[0])
    return a.numel() == b.numel() and bool((a == b).all())


def coalesce(edge_index, edge_attrs, num_nodes=None):
    """``torch_geometric.utils.coalesce(reduce='add')``: sort by row*N+col and sum
    the attributes of duplicate entries."""
    n = maybe_num_nodes(edge_index, num_nodes)
    key = edge_index[0] * n + edge_index[1]
    key, perm = torch.sort(key, stable=True)
    edge_index = edge_index[:, perm]
    edge_attrs = [a[perm] for a in edge_attrs]
    first = torch.ones_like(key, dtype=torch.bool)
    first[1:] = key[1:] > key[:-1]
    if bool(first.all()):
        return edge_index, edge_attrs
    slot = torch.cumsum(first.to(torch.int64), 0) - 1
    out_attrs = []
    for a in edge_attrs:
        o = torch.zeros((int(first.sum()),) + tuple(a.shape[1:]), dtype=a.dtype, device=a.device)
        o.index_add_(0, slot, a)
        out_attrs.append(o)
    return edge_index[:, first], out_attrs


def to_undirected(edge_index, edge_attrs, num_nodes=None):
    """``to_undirected(edge_index, [a, b])`` as called at ``delete_gnn.py:175``:
    concatenate the flipped list (attributes duplicated) then coalesce."""
    row, col = edge_index
    both = torch.stack([torch.cat([row, col]), torch.cat([col, row])], 0)
    attrs = [torch.cat([a, a], 0) for a in edge_attrs]
    return coalesce(both, attrs, num_nodes)


def k_hop_subgraph(node_idx, num_hops, edge_index, num_nodes=None):
    """``k_hop_subgraph(..., relabel_nodes=False, flow='source_to_target')`` as
    called at ``delete_gnn.py:128-140`` and ``prepare_dataset.py:202-206``.

    With this flow ``row = edge_index[1]`` (targets) and ``col = edge_index[0]``
    (sources): each hop marks the last frontier, selects the edges whose TARGET is
    marked and adds their SOURCES.  On the reference's directed ``row<col`` lists
    that walks only to lower-index neighbours."""
    n = maybe_num_nodes(edge_index, num_nodes)
    col, row = edge_index[0], edge_index[1]
    node_idx = torch.as_tensor(node_idx, dtype=torch.int64, device=edge_index.device).flatten()
    subsets = [node_idx]
    node_mask = torch.zeros(n, dtype=torch.bool, device=edge_index.device)
    for _ in range(num_hops):
        node_mask.fill_(False)
        node_mask[subsets[-1]] = True
        edge_mask = node_mask[row]
        subsets.append(col[edge_mask])
    subset, inv = torch.cat(subsets).unique(return_inverse=True)
    inv = inv[:node_idx.numel()]
    node_mask.fill_(False)
    node_mask[subset] = True
    edge_mask = node_mask[row] & node_mask[col]
    return subset, edge_index[:, edge_mask], inv, edge_mask


# ----------------------------------------------------------------- message passing
def _scatter_add(msg, index, num_nodes):
    out = torch.zeros((num_nodes,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
    return out.index_add_(0, index, msg)


def gcn_norm(edge_index, num_nodes, dtype):
    """``gcn_norm(improved=False, add_self_loops=True)``; recomputed on every
    GCNConv call because the reference leaves ``cached=False`` (gcn.py:11-12)."""
    ei = add_remaining_self_loops(edge_index, num_nodes)
    w = torch.ones(ei.size(1), dtype=dtype, device=ei.device)
    src, dst = ei[0], ei[1]
    deg = torch.zeros(num_nodes, dtype=dtype, device=ei.device).index_add_(0, dst, w)
    dinv = deg.pow(-0.5)
    dinv.masked_fill_(dinv == float('inf'), 0)
    return ei, dinv[src] * w * dinv[dst]


def gcn_conv(x, edge_index, weight, bias):
    """GCNConv(in,out) defaults (gcn.py:11-12): out = A_hat (x W^T) + b."""
    n = x.size(0)
    ei, norm = gcn_norm(edge_index, n, x.dtype)
    h = x @ weight.t()
    msg = norm.view(-1, 1) * h.index_select(0, ei[0])
    out = _scatter_add(msg, ei[1], n)
    if bias is not None:
        out = out + bias
    return out


def segment_softmax(e, index, num_nodes):
    """``torch_geometric.utils.softmax`` grouped by target index."""
    emax = torch.full((num_nodes,) + tuple(e.shape[1:]), float('-inf'), dtype=e.dtype, device=e.device)
    emax = emax.scatter_reduce(0, index.view(-1, *([1] * (e.dim() - 1))).expand_as(e), e.detach(),
                               reduce='amax', include_self=True)
    out = (e - emax.index_select(0, index)).exp()
    denom = _scatter_add(out, index, num_nodes) + 1e-16
    return out / denom.index_select(0, index)


def gat_conv(x, edge_index, weight, att_src, att_dst, bias, negative_slope=0.2):
    """GATConv(in,out) defaults heads=1, concat=True, dropout=0 (gat.py:11-12)."""
    n = x.size(0)
    c = weight.size(0)
    h = (x @ weight.t()).view(n, 1, c)
    a_s = (h * att_src).sum(-1)          # [n,1]
    a_d = (h * att_dst).sum(-1)
    ei = add_remaining_self_loops(edge_index, n)
    src, dst = ei[0], ei[1]
    e = F.leaky_relu(a_s.index_select(0, src) + a_d.index_select(0, dst), negative_slope)
    alpha = segment_softmax(e, dst, n)   # [nnz,1]
    msg = h.index_select(0, src) * alpha.unsqueeze(-1)
    out = _scatter_add(msg, dst, n).view(n, c)
    if bias is not None:
        out = out + bias
    return out


def gin_conv(x, edge_index, lin_weight, lin_bias, eps=0.0):
    """GINConv(nn.Linear(in,out)), eps=0 buffer (gin.py:11-12): aggregation runs at
    the INPUT width, no self-loop insertion, no normalisation."""
    n = x.size(0)
    agg = _scatter_add(x.index_select(0, edge_index[0]), edge_index[1], n)
    return F.linear(agg + (1.0 + eps) * x, lin_weight, lin_bias)


def rgcn_conv(x, edge_index, edge_type, weight, root, bias):
    """RGCNConv(in,out,R,num_blocks=B|None), aggr='mean' (rgcn.py:17-22), float
    input.  ``weight`` is [R,in,out] or block-diagonal [R,B,in/B,out/B].  The
    per-relation Python loop is PyG's own structure and is kept because this
    function is also the CPU timing baseline."""
    n = x.size(0)
    num_rel = weight.size(0)
    out_dim = root.size(1)
    out = torch.zeros(n, out_dim, dtype=x.dtype, device=x.device)
    for r in range(num_rel):
        sel = edge_type == r
        src, dst = edge_index[0][sel], edge_index[1][sel]
        s = _scatter_add(x.index_select(0, src), dst, n)
        cnt = torch.zeros(n, dtype=x.dtype, device=x.device).index_add_(0, dst, torch.ones(dst.numel(), dtype=x.dtype, device=x.device))
        m = s / cnt.clamp(min=1).view(-1, 1)
        if weight.dim() == 4:
            m = m.view(n, weight.size(1), weight.size(2))
            out = out + torch.einsum('abc,bcd->abd', m, weight[r]).contiguous().view(n, out_dim)
        else:
            out = out + m @ weight[r]
    out = out + x @ root
    if bias is not None:
        out = out + bias
    return out
