"""Generates tests/golden/oracle_small.npz from the CPU oracle (fp64) on a seeded 2%-scale
Cora-shaped case:   python -m tests.golden.make_golden [nodeemb]   (`nodeemb`: only oracle_nodeemb_small.npz)
The reference has no golden vectors and cannot be imported here (PyG missing), so these
vectors pin the ORACLE, not the reference: PARITY UNPINNED (oracle/__init__.py)."""
import os

import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_small.npz')
OUT_NODEEMB = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_nodeemb_small.npz')


def compute():
    from oracle import unlearn as OU
    from tests import util as U
    out = {}
    shape, raw, df, data, neg = U.make_case('cora', 0.02)
    out['train_pos_edge_index'] = data.train_pos_edge_index.numpy()
    for k in ('sdf_mask', 'df_mask', 'sdf_node_1hop_mask', 'sdf_node_2hop_mask'):
        out[k] = data[k].numpy()
    out['neg'] = neg.numpy()
    d64 = data.clone()
    d64.x = data.x.double()
    for gnn in ('gcn', 'gat', 'gin'):
        om = U.oracle_model(gnn, shape, data, dtype=torch.float64)
        with torch.no_grad():
            zo = om.get_original_embeddings(d64.x, d64.train_pos_edge_index[:, d64.dr_mask])
        loss, lr, ll, z = OU.edge_form_loss(om, d64, neg, zo)
        loss.backward()
        out[f'{gnn}_z'] = z.detach().numpy()
        out[f'{gnn}_losses'] = np.array([float(loss), float(lr), float(ll)])
        out[f'{gnn}_dW1'] = om.deletion1.deletion_weight.grad.numpy()
        out[f'{gnn}_dW2'] = om.deletion2.deletion_weight.grad.numpy()
    return out


def compute_nodeemb():
    """Loss curves and Del weights after 3 epochs of the oracle's node-embedding epoch
    (oracle.unlearn.nodeemb_epoch, gnndelete_nodeemb.py:191-299), every loss_type, alpha = 0.4."""
    from tests.test_nodeemb_cpu import _oracle_run
    out = {}
    for lt in ('both_all', 'both_layerwise', 'only2_layerwise', 'only2_all', 'only1'):
        om, _, hist = _oracle_run(lt)
        out[f'{lt}_hist'] = hist.numpy()
        out[f'{lt}_W1'] = om.deletion1.deletion_weight.detach().numpy()[::4]      # every 4th row keeps the fixture small
        out[f'{lt}_W2'] = om.deletion2.deletion_weight.detach().numpy()[::4]
    return out


if __name__ == '__main__':
    import sys
    if 'nodeemb' not in sys.argv[1:]:
        np.savez_compressed(OUT, **compute())
        print(OUT, os.path.getsize(OUT))
    np.savez_compressed(OUT_NODEEMB, **compute_nodeemb())
    print(OUT_NODEEMB, os.path.getsize(OUT_NODEEMB))
