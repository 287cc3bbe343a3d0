"""Seeded synthetic inputs with the field names / dtypes the reference consumes.

The reference reads ``(dataset, data)`` from ``d_{seed}.pkl`` (written by
``prepare_dataset.py:194-195``) where ``data.train_pos_edge_index`` is a
*directed* ``row < col`` int64 ``[2, E]`` list (``prepare_dataset.py:99``) and
``df_{seed}.pt`` holds Df candidate masks.  There is no network here, so every
BASELINE.json config is reproduced as a Chung-Lu power-law graph of the same
node / edge / feature shape.

Nothing here is on the timed path; it only manufactures inputs that are handed
identically to the CUDA implementation and to the CPU oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace

import torch

from .data import GraphData


@dataclass(frozen=True)
class Shape:
    name: str
    gnn: str
    num_nodes: int
    num_edges: int          # columns of the directed row<col train list
    num_deleted: int        # sampled directed Df edges
    in_dim: int = 128
    hidden_dim: int = 128
    out_dim: int = 64
    exponent: float = 0.6
    num_edge_type: int = 0  # >0 => knowledge graph (head->tail triples)
    capped: bool = False    # clamp expected degrees at sqrt(2E) (config 5 only)

    def scaled(self, factor: float) -> 'Shape':
        """Same degree law at ``factor`` x the node/edge/deletion counts."""
        if factor == 1.0:
            return self
        return replace(
            self,
            name=f'{self.name}@{factor:g}',
            num_nodes=max(16, int(self.num_nodes * factor)),
            num_edges=max(32, int(self.num_edges * factor)),
            num_deleted=max(2, int(self.num_deleted * factor)),
        )


# BASELINE.json configs (README.md:44-52 shapes)
SHAPES = {
    'cora': Shape('cora', 'gcn', 19_793, 126_842, 6_342),
    'pubmed': Shape('pubmed', 'gat', 19_717, 88_648, 4_432, in_dim=500),
    'collab': Shape('collab', 'gcn', 235_368, 1_285_465, 117_905),
    'biokg': Shape('biokg', 'rgcn', 93_773, 5_088_434, 127_210, num_edge_type=51),
    'powerlaw10m': Shape('powerlaw10m', 'gcn', 10_000_000, 200_000_000, 10_000_000, exponent=0.5,
                         capped=True),
}


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def chung_lu_edges(num_nodes: int, num_edges: int, exponent: float, seed: int,
                   device='cpu', directed_kg: bool = False, capped: bool = False) -> torch.Tensor:
    """``[2, num_edges]`` int64, duplicate-free, no self loops.

    Expected degree of node i is proportional to ``(rank_i + 1) ** -exponent``
    where ``rank`` is a seeded random permutation, so node ids are *not* sorted
    by degree.  Homogeneous graphs are returned in the reference's on-disk
    convention ``row < col``; KG triples keep the sampled head->tail direction.
    """
    device = torch.device(device)
    g = _gen(seed, device)
    n = num_nodes
    max_pairs = n * (n - 1) // 2
    if num_edges > max_pairs:
        raise ValueError(f'{num_edges} edges do not fit {n} nodes')
    w = (torch.arange(n, device=device, dtype=torch.float64) + 1.0).pow(-exponent)
    if capped:
        # cap the heaviest expected degree at sqrt(2E) so the Chung-Lu model stays
        # simple-graph-feasible on the 10M-node config (SURVEY.md §8(d) "capped")
        cap = math.sqrt(2.0 * num_edges) * w.sum().item() / (2.0 * num_edges)
        w = w.clamp(max=cap)
    cdf = torch.cumsum(w, 0)
    cdf = cdf / cdf[-1]
    relabel = torch.randperm(n, generator=g, device=device)

    keys = torch.empty(0, dtype=torch.int64, device=device)
    need = num_edges
    while keys.numel() < num_edges:
        m = int(need * 1.25) + 1024
        a = torch.searchsorted(cdf, torch.rand(m, generator=g, device=device, dtype=torch.float64))
        b = torch.searchsorted(cdf, torch.rand(m, generator=g, device=device, dtype=torch.float64))
        a = relabel[a.clamp_(max=n - 1)]
        b = relabel[b.clamp_(max=n - 1)]
        keep = a != b
        a, b = a[keep], b[keep]
        if directed_kg:
            new = a * n + b
        else:
            new = torch.minimum(a, b) * n + torch.maximum(a, b)
        keys = torch.unique(torch.cat([keys, new]))
        need = num_edges - keys.numel()
    pick = torch.randperm(keys.numel(), generator=g, device=device)[:num_edges]
    keys = keys[pick]
    return torch.stack([keys // n, keys % n], 0)


def make_graph(shape, seed: int = 42, device='cpu', with_eval_edges: bool = True) -> GraphData:
    """The equivalent of loading ``d_{seed}.pkl``: a directed train split (+ small
    val/test splits with sampled negatives, ``prepare_dataset.py:31-136``)."""
    if isinstance(shape, str):
        shape = SHAPES[shape]
    device = torch.device(device)
    kg = shape.num_edge_type > 0
    n_eval = max(2, shape.num_edges // 18) if with_eval_edges else 0   # 5 / 90 of train
    total = shape.num_edges + 2 * n_eval
    ei = chung_lu_edges(shape.num_nodes, total, shape.exponent, seed, device, directed_kg=kg,
                        capped=shape.capped)
    g = _gen(seed, device)
    data = GraphData(num_nodes=shape.num_nodes)
    data.train_pos_edge_index = ei[:, :shape.num_edges].contiguous()
    if kg:
        gt = _gen(seed + 2, device)
        et = torch.randint(0, shape.num_edge_type, (total,), generator=gt, device=device)
        data.train_edge_type = et[:shape.num_edges].contiguous()
        data.x = torch.arange(shape.num_nodes, device=device)
    else:
        data.x = torch.randn(shape.num_nodes, shape.in_dim, generator=g, device=device)
    if with_eval_edges:
        for i, stage in enumerate(('val', 'test')):
            lo = shape.num_edges + i * n_eval
            data[f'{stage}_pos_edge_index'] = ei[:, lo:lo + n_eval].contiguous()
            data[f'{stage}_neg_edge_index'] = torch.randint(
                0, shape.num_nodes, (2, n_eval), generator=g, device=device)
            if kg:
                data[f'{stage}_edge_type'] = et[lo:lo + n_eval].contiguous()
    return data


def sample_df_mask(num_edges: int, num_deleted: int, seed: int = 42, device='cpu',
                   candidates: torch.Tensor | None = None) -> torch.Tensor:
    """Directed Df mask, following ``delete_gnn.py:95-110``: a random
    ``num_deleted``-subset of the candidate columns (all columns when no
    ``df_{seed}.pt`` candidate mask is supplied)."""
    device = torch.device(device)
    g = _gen(seed, device)
    if candidates is None:
        pool = torch.arange(num_edges, device=device)
    else:
        pool = candidates.nonzero().squeeze(1)
    idx = pool[torch.randperm(pool.numel(), generator=g, device=device)[:num_deleted]]
    mask = torch.zeros(num_edges, dtype=torch.bool, device=device)
    mask[idx] = True
    return mask


def supplied_negatives(num_nodes: int, count: int, seed: int = 43, device='cpu') -> torch.Tensor:
    """``negative_sampling`` is not reproducible across implementations
    (SURVEY.md §9.7); parity and bench runs hand the same ``[2, count]`` set to
    both sides."""
    device = torch.device(device)
    return torch.randint(0, num_nodes, (2, count), generator=_gen(seed, device), device=device)
