"""BASELINE configs 1, 2 and 4 at their FULL node / edge / feature shapes against the fp64 CPU oracle (1e-5): the same
comparisons the small-scale tests make (tests/test_gpu_engine.py, tests/test_gpu_gat_rgcn.py), on the graphs the
configs name.  Config 3 (Collab) is tests/test_gpu_fullsize.py; config 5's partition is tests/test_gpu_dist.py.
The oracle side of the two citation graphs takes seconds, the BioKG step (per-relation Python loop over 10 M
message-passing edges, gnndelete_nodeemb.py:744-798) about a minute of host time."""
import pytest

pytestmark = pytest.mark.gpu


def test_config1_cora_gcndelete_dense_ni_full_size(lib):
    """19,793 nodes / 126,842 directed edges: GCNDelete epoch with train_fullbatch's dense S2 x S2 NI
    (gnndelete.py:163-193, 239-241) - losses, both Del gradients, CUDA-graph replay."""
    from tests.test_gpu_engine import dense_ni_case
    dense_ni_case(1.0)


def test_config1_cora_gcndelete_edge_form_full_size(lib):
    """Same graph, the edge-form NI of train_minibatch (gnndelete.py:347-409) through the fused engine."""
    import torch
    from gnndelete_b200 import models as M
    from gnndelete_b200.engine import GCNDeleteEngine
    from oracle import unlearn as OU
    from tests import util as U
    shape, raw, df, data, neg = U.make_case('cora', 1.0)
    om = U.oracle_model('gcn', shape, data, dtype=torch.float64)
    d64 = data.clone(); d64.x = data.x.double()
    with torch.no_grad():
        zo = om.get_original_embeddings(d64.x, d64.train_pos_edge_index[:, d64.dr_mask])
    loss, lr, ll, _ = OU.edge_form_loss(om, d64, neg, zo)
    loss.backward()
    m = M.GCNDelete(U.args_for(shape), data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
    m.load_state_dict({k: v.float() for k, v in om.state_dict().items()})
    m = m.to('cuda')
    eng = GCNDeleteEngine(m, data.clone().to('cuda'), neg.to('cuda'), z_ori=zo.float().to('cuda'), hoist_layer1=False)
    U.assert_close(eng.forward_backward(), torch.stack([loss, lr, ll]).detach(), what='cora losses')
    U.assert_close(eng.params[0].grad, om.deletion1.deletion_weight.grad, what='cora dW_del1')
    U.assert_close(eng.params[1].grad, om.deletion2.deletion_weight.grad, what='cora dW_del2')


def test_config2_pubmed_gatdelete_full_size(lib):
    """19,717 nodes / 88,648 directed edges, F_in = 500 (CitationFull-PubMed): GATDelete embeddings of both layers,
    losses and both Del gradients (gat.py:11-12 edge-softmax aggregation forward and backward)."""
    from tests.test_gpu_gat_rgcn import gat_delete_case
    gat_delete_case(500, 1.0)


def test_config4_biokg_rgcndelete_full_size(lib):
    """93,773 nodes / 5,088,434 triples / 51 relation types (R = 102, 4 blocks): RGCNDelete embeddings of both
    layers, the KG node-embedding step's losses (gnndelete_nodeemb.py:744-798), both Del gradients, DistMult."""
    from tests.test_gpu_gat_rgcn import rgcn_delete_case
    rgcn_delete_case(51, 1.0)
