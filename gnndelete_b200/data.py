"""Minimal attribute container standing in for ``torch_geometric.data.Data``.

The reference threads a PyG ``Data`` object through ``delete_gnn.py`` and the
trainers and only ever uses attribute access, ``data[key]`` lookup
(reference ``framework/trainer/base.py:232-233``), ``hasattr`` and ``.to(device)``
(``framework/trainer/gnndelete.py:140``).  PyG is not a dependency of this
package, so this container offers exactly that surface.
"""
from __future__ import annotations

import torch


class GraphData:
    def __init__(self, **fields):
        for k, v in fields.items():
            setattr(self, k, v)

    # data['val_pos_edge_index'] style access (base.py:232)
    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return hasattr(self, key)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith('_')]

    def to(self, device, non_blocking=False):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self

    def cpu(self):
        return self.to('cpu')

    def clone(self):
        out = GraphData()
        for k in self.keys():
            v = getattr(self, k)
            setattr(out, k, v.clone() if torch.is_tensor(v) else v)
        return out

    def __repr__(self):
        parts = []
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                parts.append(f'{k}={list(v.shape)}')
            else:
                parts.append(f'{k}={v!r}')
        return 'GraphData(' + ', '.join(parts) + ')'
