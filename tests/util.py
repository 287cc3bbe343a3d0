"""Shared helpers for the parity tests: seeded small cases handed identically to the
CUDA path and to the CPU oracle."""
import dataclasses
import types

import torch

from gnndelete_b200 import synthetic as S
from oracle import models as OM
from oracle import unlearn as OU

TOL_FP32 = 1e-5      # north_star: embeddings, logits, losses, Del gradients within 1e-5 relative (fp32)


def args_for(shape):
    return types.SimpleNamespace(in_dim=shape.in_dim, hidden_dim=shape.hidden_dim, out_dim=shape.out_dim)


def make_case(name='cora', scale=0.05, seed=42, in_dim=None, gnn=None):
    """(shape, raw data, unlearning data, negatives) on the CPU."""
    shape = S.SHAPES[name].scaled(scale)
    if in_dim is not None:
        shape = dataclasses.replace(shape, in_dim=in_dim)
    raw = S.make_graph(shape, seed=seed)
    df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=seed)
    kg = shape.num_edge_type if shape.num_edge_type > 0 else None
    data = OU.build_unlearning_data(raw, df, num_edge_type=kg)
    neg = S.supplied_negatives(shape.num_nodes, int(data.df_mask.sum()), seed=seed + 1)
    return shape, raw, df, data, neg


def randomize(model, seed=0, del_scale=0.05):
    """Non-degenerate parameters: random biases and Del weights (the reference's init is
    zero bias / ones/1000, which hides bias and Del-weight indexing mistakes)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('bias'):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            if 'deletion_weight' in n:
                p.copy_(torch.eye(p.shape[0]) + torch.randn(p.shape, generator=g) * del_scale)
    return model


def oracle_model(gnn, shape, data, dtype=torch.float32, seed=0, delete=True, **kw):
    torch.manual_seed(seed)
    cls = (OM.DELETE_MODELS if delete else OM.MODELS)[gnn]
    if delete:
        m = cls(args_for(shape), mask_1hop=data.sdf_node_1hop_mask, mask_2hop=data.sdf_node_2hop_mask, **kw)
    else:
        m = cls(args_for(shape), **kw)
    randomize(m, seed)
    return m.to(dtype)


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    scale = b.abs().max().item()
    return (a - b).abs().max().item() / (scale if scale > 0 else 1.0)


def assert_close(a, b, tol=TOL_FP32, what=''):
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    e = rel_err(a, b)
    assert e <= tol, f'{what}: relative error {e:.3e} > {tol:.1e}'
    return e
