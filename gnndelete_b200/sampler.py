"""``GraphSAINTRandomWalkSampler`` as the reference's mini-batch loops use it
(``framework/trainer/gnndelete.py:333-337``: ``batch_size=args.batch_size, walk_length=2, num_steps=args.num_steps``,
default ``sample_coverage=0`` so no normalisation statistics), re-stated with tensor ops so that it runs on the
device next to the graph instead of in DataLoader worker processes.

PyG semantics kept (2.0 - 2.2, ``torch_geometric/loader/graph_saint.py`` + torch_sparse ``random_walk`` /
``saint_subgraph``): one batch = ``batch_size`` uniformly random start nodes, a ``walk_length``-step uniform random
walk from each along ``edge_index[0] -> edge_index[1]`` (a node without out-edges stays where it is), the sorted unique
set of visited nodes, and the subgraph INDUCED on it with nodes relabelled in that order and edges in (row, col) order.
Every tensor attribute whose first dimension is ``num_nodes`` is sliced by the node set, every one whose first
dimension is ``num_edges`` by the kept edge ids (node-sized wins if both match), everything else is passed through.
The random draws differ from PyG's (different generators), like every sampler reimplementation's."""
from __future__ import annotations

import torch

from .data import GraphData


class GraphSAINTRandomWalkSampler:
    def __init__(self, data, batch_size, walk_length, num_steps=1, generator=None):
        ei = data.edge_index
        self.data = data
        self.N = int(data.num_nodes)
        self.E = ei.shape[1]
        self.batch_size, self.walk_length, self.num_steps = int(batch_size), int(walk_length), int(num_steps)
        self.generator = generator
        dev = ei.device
        # adjacency sorted by (row, col), remembering the original edge ids (SparseTensor(row, col, value=arange(E)))
        order = torch.argsort(ei[0] * self.N + ei[1], stable=True)
        self.row, self.col, self.edge_id = ei[0][order], ei[1][order], order
        deg = torch.bincount(self.row, minlength=self.N)
        self.rowptr = torch.zeros(self.N + 1, dtype=torch.int64, device=dev)
        torch.cumsum(deg, 0, out=self.rowptr[1:])
        self.deg = deg

    def __len__(self):
        return self.num_steps

    def random_walk(self, start):
        """``[len(start), walk_length + 1]`` node ids (torch_sparse ``random_walk``)."""
        walk = [start]
        cur = start
        for _ in range(self.walk_length):
            d = self.deg[cur]
            r = torch.rand(cur.shape, generator=self.generator, device=cur.device)
            pick = torch.minimum((r * d).long(), (d - 1).clamp(min=0))
            nxt = self.col[(self.rowptr[cur] + pick).clamp(max=max(self.E - 1, 0))] if self.E else cur
            cur = torch.where(d > 0, nxt, cur)
            walk.append(cur)
        return torch.stack(walk, 1)

    def sample_nodes(self):
        dev = self.row.device
        start = torch.randint(0, self.N, (self.batch_size,), generator=self.generator, device=dev)
        return self.random_walk(start).reshape(-1)

    def subgraph(self, node_idx):
        """(relabelled ``edge_index``, kept original edge ids) of the subgraph induced on the sorted ``node_idx``."""
        dev = node_idx.device
        local = torch.full((self.N,), -1, dtype=torch.int64, device=dev)
        local[node_idx] = torch.arange(node_idx.numel(), device=dev)
        keep = (local[self.row] >= 0) & (local[self.col] >= 0)
        return torch.stack([local[self.row[keep]], local[self.col[keep]]]), self.edge_id[keep]

    def __iter__(self):
        for _ in range(self.num_steps):
            node_idx = torch.unique(self.sample_nodes())
            edge_index, edge_idx = self.subgraph(node_idx)
            batch = GraphData(num_nodes=int(node_idx.numel()), edge_index=edge_index)
            for key in self.data.keys():
                if key in ('edge_index', 'num_nodes'):
                    continue
                item = self.data[key]
                if torch.is_tensor(item) and item.dim() > 0 and item.size(0) == self.N:
                    batch[key] = item[node_idx]
                elif torch.is_tensor(item) and item.dim() > 0 and item.size(0) == self.E:
                    batch[key] = item[edge_idx]
                else:
                    batch[key] = item
            yield batch
