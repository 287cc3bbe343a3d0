"""GPU parity against the committed golden vectors (tests/golden/oracle_small.npz)."""
import os

import numpy as np
import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu
DEV = 'cuda'
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'oracle_small.npz')


@pytest.mark.parametrize('gnn', ['gcn', 'gat', 'gin'])
def test_models_against_golden(lib, gnn):
    from gnndelete_b200 import masks as MK
    from gnndelete_b200 import models as M
    from gnndelete_b200.losses import EdgeLossPlan
    gold = np.load(GOLDEN)
    shape, raw, df, data, neg = U.make_case('cora', 0.02)
    assert np.array_equal(neg.numpy(), gold['neg'])
    # masks: bit-exact (CUDA mask pipeline vs golden)
    dd = MK.build_unlearning_data(raw.clone().to(DEV), df.to(DEV))
    for k in ('train_pos_edge_index', 'sdf_mask', 'df_mask', 'sdf_node_1hop_mask', 'sdf_node_2hop_mask'):
        assert np.array_equal(dd[k].cpu().numpy(), gold[k]), k
    om = U.oracle_model(gnn, shape, data, dtype=torch.float32)       # same seeded weights as the generator
    cls = {'gcn': M.GCNDelete, 'gat': M.GATDelete, 'gin': M.GINDelete}[gnn]
    m = cls(U.args_for(shape), dd.sdf_node_1hop_mask, dd.sdf_node_2hop_mask)
    m.load_state_dict(om.state_dict())
    m = m.to(DEV)
    ei = dd.train_pos_edge_index
    with torch.no_grad():
        zo = m.get_original_embeddings(dd.x, ei[:, dd.dr_mask])
    z = m(dd.x, ei[:, dd.sdf_mask])
    sdf = ei[:, dd.sdf_mask]
    plan = EdgeLossPlan(ei[:, dd.df_mask], neg.to(DEV), sdf[:, sdf[0] < sdf[1]], dd.num_nodes, z_ori=zo)
    from gnndelete_b200.losses import edge_loss
    loss, lr, ll = edge_loss(z, plan)
    loss.backward()
    U.assert_close(z, torch.from_numpy(gold[f'{gnn}_z']), what='z')
    U.assert_close(torch.stack([loss, lr, ll]), torch.from_numpy(gold[f'{gnn}_losses']), what='losses')
    U.assert_close(m.deletion1.deletion_weight.grad, torch.from_numpy(gold[f'{gnn}_dW1']), what='dW1')
    U.assert_close(m.deletion2.deletion_weight.grad, torch.from_numpy(gold[f'{gnn}_dW2']), what='dW2')
