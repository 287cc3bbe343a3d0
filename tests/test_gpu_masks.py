"""GPU parity (bit-exact): k-hop deletion masks, to_undirected and the delete_gnn.py mask pipeline."""
import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('name,scale', [('cora', 0.05), ('cora', 0.5), ('pubmed', 0.2)])
def test_mask_pipeline_bit_exact(lib, name, scale):
    from gnndelete_b200 import masks as MK
    shape, raw, df, data, neg = U.make_case(name, scale)
    out = MK.build_unlearning_data(raw.clone().to(DEV), df.to(DEV))
    for k in ['train_pos_edge_index', 'sdf_mask', 'df_mask', 'dr_mask', 'sdf_node_1hop_mask',
              'sdf_node_2hop_mask', 'directed_df_edge_index']:
        assert torch.equal(out[k].cpu(), data[k]), k
    # delete_gnn.py:147-148 asserts
    assert int(out.sdf_node_2hop_mask.sum()) >= int(out.sdf_node_1hop_mask.sum()) > 0


def test_khop_directed_quirk(lib):
    """On a directed row<col path 0-1-2-3 a hop from node 3 walks only to lower ids."""
    from gnndelete_b200 import masks as MK
    from oracle import pyg_ops as P
    ei = torch.tensor([[0, 1, 2, 4], [1, 2, 3, 5]])
    for seeds, hops in [([3], 1), ([3], 2), ([0], 2), ([1, 5], 1), ([], 1)]:
        s_o, e_o, inv_o, m_o = P.k_hop_subgraph(torch.tensor(seeds, dtype=torch.long), hops, ei, num_nodes=6)
        s_g, e_g, inv_g, m_g = MK.k_hop_subgraph(torch.tensor(seeds, dtype=torch.long), hops, ei.to(DEV), num_nodes=6)
        assert torch.equal(m_g.cpu(), m_o), (seeds, hops)
        assert torch.equal(e_g.cpu(), e_o)
        assert torch.equal(s_g.cpu(), s_o)
        assert torch.equal(inv_g.cpu(), inv_o)


def test_to_undirected_merges_duplicates(lib):
    from gnndelete_b200 import masks as MK
    from oracle import pyg_ops as P
    # (1,2) and (2,1) both present -> duplicates after symmetrisation, attributes add up
    ei = torch.tensor([[1, 2, 0, 3, 3], [2, 1, 4, 3, 0]])
    a = torch.tensor([1, 0, 1, 1, 0], dtype=torch.int32)
    b = torch.tensor([5, 7, 0, 2, 1], dtype=torch.int32)
    sym_o, (a_o, b_o) = P.to_undirected(ei, [a, b])
    sym_g, (a_g, b_g) = MK.to_undirected(ei.to(DEV), [a.to(DEV), b.to(DEV)])
    assert torch.equal(sym_g.cpu(), sym_o)
    assert torch.equal(a_g.cpu(), a_o) and torch.equal(b_g.cpu(), b_o)
    assert MK.to_undirected(torch.empty(2, 0, dtype=torch.long, device=DEV)).shape == (2, 0)


def test_kg_pipeline_bit_exact(lib):
    from gnndelete_b200 import masks as MK
    shape, raw, df, data, neg = U.make_case('biokg', 0.002)
    out = MK.build_unlearning_data(raw.clone().to(DEV), df.to(DEV), num_edge_type=shape.num_edge_type)
    for k in ['edge_index', 'edge_type', 'sdf_mask', 'df_mask', 'dr_mask', 'sdf_node_1hop_mask',
              'sdf_node_2hop_mask', 'directed_df_edge_index', 'directed_df_edge_type']:
        assert torch.equal(out[k].cpu(), data[k]), k
