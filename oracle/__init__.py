"""TEST INFRASTRUCTURE ONLY.

CPU restatement (eager PyTorch, fp32/fp64) of the arithmetic GNNDelete's
unlearning hot path performs: the PyTorch-Geometric operators it calls
(``pyg_ops``), the reference's own model layer (``models``) and its unlearning
driver / loss bodies (``unlearn``).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it, and
there only as the checker or the timed CPU baseline — the product package
``gnndelete_b200`` never does.

PARITY UNPINNED: the reference ships no tests, golden vectors or stored
tensors, does not import as shipped (``framework/__init__.py:10`` imports a
missing module) and delegates the arithmetic to an unpinned, uninstalled
PyTorch-Geometric (2.0.3 … 2.2.x by API usage).  The oracle's authority rests on
line-by-line correspondence with the cited in-tree files, PyG's published
default-path semantics, float64 ``gradcheck``, an independent
dense-matrix / networkx cross-check, and the known-answer vectors of PyG's own docstrings
and unit tests (``k_hop_subgraph``, ``to_undirected``, ``softmax``, ``add_remaining_self_loops``;
written down from the published sources, not executed against an install) in
``tests/test_oracle.py``.
"""
