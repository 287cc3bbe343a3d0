"""TEST INFRASTRUCTURE — CPU restatement of the reference's model layer
(``framework/models/{gcn,gat,gin,rgcn,deletion}.py``) on top of ``pyg_ops``.
PARITY UNPINNED (see ``oracle/pyg_ops.py``).  Never imported by the product.

State-dict keys follow the PyG layers the reference instantiates so the same
checkpoint loads into the oracle and into the CUDA models.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pyg_ops as P


def _glorot(t):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)
    return t


class _Lin(nn.Module):          # PyG ``Linear(bias=False)`` -> key ``lin.weight``
    def __init__(self, i, o):
        super().__init__()
        self.weight = nn.Parameter(_glorot(torch.empty(o, i)))


class GCNConv(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.lin = _Lin(i, o)
        self.bias = nn.Parameter(torch.zeros(o))

    def forward(self, x, edge_index):
        return P.gcn_conv(x, edge_index, self.lin.weight, self.bias)


class GATConv(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.lin_src = _Lin(i, o)
        self.lin_dst = self.lin_src
        self.att_src = nn.Parameter(_glorot(torch.empty(1, 1, o)))
        self.att_dst = nn.Parameter(_glorot(torch.empty(1, 1, o)))
        self.bias = nn.Parameter(torch.zeros(o))

    def forward(self, x, edge_index):
        return P.gat_conv(x, edge_index, self.lin_src.weight, self.att_src, self.att_dst, self.bias)


class GINConv(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.nn = nn.Linear(i, o)
        self.register_buffer('eps', torch.tensor([0.0]))

    def forward(self, x, edge_index):
        return P.gin_conv(x, edge_index, self.nn.weight, self.nn.bias, float(self.eps))


class RGCNConv(nn.Module):
    def __init__(self, i, o, num_relations, num_blocks=None):
        super().__init__()
        if num_blocks is None:
            self.weight = nn.Parameter(_glorot(torch.empty(num_relations, i, o)))
        else:
            self.weight = nn.Parameter(_glorot(torch.empty(num_relations, num_blocks, i // num_blocks, o // num_blocks)))
        self.root = nn.Parameter(_glorot(torch.empty(i, o)))
        self.bias = nn.Parameter(torch.zeros(o))

    def forward(self, x, edge_index, edge_type):
        return P.rgcn_conv(x, edge_index, edge_type, self.weight, self.root, self.bias)


# ------------------------------------------------------------------ base encoders
# Test hook: a *Delete model whose attribute ``relu_mask_override`` holds a bool tensor uses it instead of the sign test of
# the inter-layer ReLU (``relu(x) := x * mask``) in its Delete forward.  A parity test sets it to the mask the
# implementation under test produced when some pre-activation lies within rounding distance of zero: there the fp64 and
# fp32 forward passes may disagree on the sign, and d relu / dx jumps by a finite amount.
def _relu(x, mask=None):
    if mask is not None:
        return x * mask.to(x.dtype)
    return F.relu(x)


class _TwoLayer(nn.Module):
    """conv1 -> ReLU -> conv2, no dropout (gcn.py:15-24, gat.py:15-24, gin.py:26-34)."""
    conv_cls = None

    def __init__(self, args, **kw):
        super().__init__()
        self.conv1 = self.conv_cls(args.in_dim, args.hidden_dim)
        self.conv2 = self.conv_cls(args.hidden_dim, args.out_dim)

    def forward(self, x, edge_index, return_all_emb=False):
        x1 = self.conv1(x, edge_index)
        x2 = self.conv2(F.relu(x1), edge_index)
        return (x1, x2) if return_all_emb else x2

    def decode(self, z, pos_edge_index, neg_edge_index=None):       # gcn.py:26-35
        ei = pos_edge_index if neg_edge_index is None else torch.cat([pos_edge_index, neg_edge_index], -1)
        return (z[ei[0]] * z[ei[1]]).sum(-1)


class GCN(_TwoLayer):
    conv_cls = GCNConv


class GAT(_TwoLayer):
    conv_cls = GATConv


class GIN(_TwoLayer):
    conv_cls = GINConv


class RGCN(nn.Module):                                               # rgcn.py:9-47
    def __init__(self, args, num_nodes, num_edge_type, **kw):
        super().__init__()
        self.num_edge_type = num_edge_type
        self.node_emb = nn.Embedding(num_nodes, args.in_dim)
        blocks = 4 if num_edge_type > 20 else None
        self.conv1 = RGCNConv(args.in_dim, args.hidden_dim, num_edge_type * 2, blocks)
        self.conv2 = RGCNConv(args.hidden_dim, args.out_dim, num_edge_type * 2, blocks)
        self.relu = nn.ReLU()
        self.W = nn.Parameter(torch.empty(num_edge_type, args.out_dim))
        nn.init.xavier_uniform_(self.W, gain=nn.init.calculate_gain('relu'))

    def forward(self, x, edge, edge_type, return_all_emb=False):
        x = self.node_emb(x)
        x1 = self.conv1(x, edge, edge_type)
        x2 = self.conv2(self.relu(x1), edge, edge_type)
        return (x1, x2) if return_all_emb else x2

    def decode(self, z, edge_index, edge_type):
        return torch.sum(z[edge_index[0]] * self.W[edge_type] * z[edge_index[1]], dim=1)


# ---------------------------------------------------------------------- Del layer
class DeletionLayer(nn.Module):                                      # deletion.py:8-29
    def __init__(self, dim, mask):
        super().__init__()
        self.dim = dim
        self.mask = mask
        self.deletion_weight = nn.Parameter(torch.ones(dim, dim) / 1000)

    def forward(self, x, mask=None):
        if mask is None:
            mask = self.mask
        if mask is None:
            return x
        new_rep = x.clone()
        new_rep[mask] = torch.matmul(new_rep[mask], self.deletion_weight)
        return new_rep


def _delete_variant(base, conv1_no_grad):
    class _Delete(base):
        def __init__(self, args, mask_1hop=None, mask_2hop=None, **kw):
            super().__init__(args, **kw)
            self.deletion1 = DeletionLayer(args.hidden_dim, mask_1hop)
            self.deletion2 = DeletionLayer(args.out_dim, mask_2hop)

        def forward(self, x, edge_index, mask_1hop=None, mask_2hop=None, return_all_emb=False):
            if conv1_no_grad:                       # deletion.py:90-91, 117-118
                with torch.no_grad():
                    x1 = self.conv1(x, edge_index)
            else:                                   # deletion.py:62-63 (GCN: no_grad commented out)
                x1 = self.conv1(x, edge_index)
            x1 = self.deletion1(x1, mask_1hop)
            x2 = self.conv2(_relu(x1, getattr(self, 'relu_mask_override', None)), edge_index)
            x2 = self.deletion2(x2, mask_2hop)
            return (x1, x2) if return_all_emb else x2

        def get_original_embeddings(self, x, edge_index, return_all_emb=False):
            return base.forward(self, x, edge_index, return_all_emb)

    _Delete.__name__ = base.__name__ + 'Delete'
    return _Delete


GCNDelete = _delete_variant(GCN, conv1_no_grad=False)
GATDelete = _delete_variant(GAT, conv1_no_grad=True)
GINDelete = _delete_variant(GIN, conv1_no_grad=True)


class RGCNDelete(RGCN):                                              # deletion.py:135-163
    def __init__(self, args, num_nodes, num_edge_type, mask_1hop=None, mask_2hop=None, **kw):
        super().__init__(args, num_nodes, num_edge_type)
        self.deletion1 = DeletionLayer(args.hidden_dim, mask_1hop)
        self.deletion2 = DeletionLayer(args.out_dim, mask_2hop)

    def forward(self, x, edge_index, edge_type, mask_1hop=None, mask_2hop=None, return_all_emb=False):
        with torch.no_grad():
            x = self.node_emb(x)
            x1 = self.conv1(x, edge_index, edge_type)
        x1 = self.deletion1(x1, mask_1hop)
        x2 = self.conv2(_relu(x1, getattr(self, 'relu_mask_override', None)), edge_index, edge_type)
        x2 = self.deletion2(x2, mask_2hop)
        return (x1, x2) if return_all_emb else x2

    def get_original_embeddings(self, x, edge_index, edge_type, return_all_emb=False):
        return RGCN.forward(self, x, edge_index, edge_type, return_all_emb)


MODELS = {'gcn': GCN, 'gat': GAT, 'gin': GIN, 'rgcn': RGCN}
DELETE_MODELS = {'gcn': GCNDelete, 'gat': GATDelete, 'gin': GINDelete, 'rgcn': RGCNDelete}
