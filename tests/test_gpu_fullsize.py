"""Full-size (BASELINE config 3, OGB-Collab shape) checks: one Del-training step against the CPU oracle at the
1e-5 bar, plus size-independent properties of the aggregation kernels on the real edge set (exact integer row
sums, linearity, bitwise reproducibility, kernel A/B agreement) and of the mask pipeline."""
import types

import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def collab(lib):
    from gnndelete_b200 import synthetic as S
    from oracle import unlearn as OU
    shape = S.SHAPES['collab']
    raw = S.make_graph(shape, seed=42, device='cpu')
    df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=42)
    data = OU.build_unlearning_data(raw, df)                      # CPU oracle pipeline
    neg = S.supplied_negatives(shape.num_nodes, int(data.df_mask.sum()), seed=43)
    return shape, raw, df, data, neg


def test_mask_pipeline_full_size_bit_exact(lib, collab):
    """k-hop masks + to_undirected on 1.29 M directed edges: CUDA pipeline == oracle, bit for bit."""
    from gnndelete_b200 import masks as MK
    shape, raw, df, data, neg = collab
    got = MK.build_unlearning_data(raw.to(DEV), df.to(DEV))
    for k in ('train_pos_edge_index', 'df_mask', 'sdf_mask', 'dr_mask', 'sdf_node_1hop_mask', 'sdf_node_2hop_mask'):
        assert torch.equal(getattr(got, k).cpu(), getattr(data, k)), k
    ei = got.train_pos_edge_index
    key = ei[0] * shape.num_nodes + ei[1]
    assert bool((key[1:] > key[:-1]).all()), 'to_undirected output must be sorted and duplicate free'
    assert bool((got.sdf_node_1hop_mask <= got.sdf_node_2hop_mask).all()), 'S1 is a subset of S2'


def test_aggregation_properties_full_size(lib, collab):
    from gnndelete_b200 import graph as G
    from gnndelete_b200 import ops
    shape, raw, df, data, neg = collab
    n = shape.num_nodes
    ei = data.train_pos_edge_index[:, data.sdf_mask].to(DEV)
    plan = G.GraphPlan(ei, n, self_loops=True, gcn_norm=True)
    csr = plan.fwd
    deg = (csr.rowptr[1:] - csr.rowptr[:-1]).float()
    gen = torch.Generator().manual_seed(5)
    for f in (64, 128):
        ones = torch.ones(n, f, device=DEV)
        out = ops.spmm(csr, ones)
        assert torch.equal(out, deg[:, None].expand(n, f)), f'row sums of ones must equal the degrees exactly (F={f})'
        x, y = torch.randn(n, f, generator=gen).to(DEV), torch.randn(n, f, generator=gen).to(DEV)
        ax, ay = ops.spmm(csr, x), ops.spmm(csr, y)
        lin = ops.spmm(csr, 0.5 * x - 2.0 * y)
        U.assert_close(lin, 0.5 * ax.double() - 2.0 * ay.double(), what=f'linearity F={f}')
        assert torch.equal(ops.spmm(csr, x), ax), 'aggregation must be bitwise reproducible'
        # the batched kernel and the row-walking kernel are independent implementations of the same sum
        was = G.BATCHED
        try:
            G.BATCHED = False
            ref = ops.spmm(csr, x, row_scale=plan.dinv, col_scale=plan.dinv)
        finally:
            G.BATCHED = was
        U.assert_close(ops.spmm(csr, x, row_scale=plan.dinv, col_scale=plan.dinv), ref, what=f'batched vs row kernel F={f}')
    # GCN normalisation: A_hat 1 has row sums dinv_i * sum_j dinv_j over the neighbourhood; symmetric => x^T A y = y^T A x
    x1, y1 = torch.randn(n, 64, generator=gen).to(DEV), torch.randn(n, 64, generator=gen).to(DEV)
    axy = (x1.double() * ops.spmm(csr, y1, row_scale=plan.dinv, col_scale=plan.dinv).double()).sum()
    ayx = (y1.double() * ops.spmm(plan.bwd, x1, row_scale=plan.dinv, col_scale=plan.dinv).double()).sum()
    assert abs(float(axy - ayx)) <= 1e-6 * abs(float(axy)) + 1e-3, 'transpose-backward is the adjoint of the forward'


def test_one_epoch_against_oracle_full_size(lib, collab):
    """Losses and both Del-weight gradients of one full-graph epoch at the Collab shape, CUDA engine (fp32,
    tcgen05 3xTF32 GEMMs) vs the CPU oracle in fp64, at the fp32 tolerance 1e-5."""
    from gnndelete_b200 import models as M
    from gnndelete_b200.engine import GCNDeleteEngine
    from oracle import unlearn as OU
    shape, raw, df, data, neg = collab
    om = U.oracle_model('gcn', shape, data, dtype=torch.float64)
    d64 = data.clone()
    d64.x = data.x.double()
    with torch.no_grad():
        zo = om.get_original_embeddings(d64.x, d64.train_pos_edge_index[:, d64.dr_mask])
    loss, lr, ll, _ = OU.edge_form_loss(om, d64, neg, zo, masks_positional=False)
    loss.backward()
    m = M.GCNDelete(U.args_for(shape), data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
    m.load_state_dict({k: v.float() for k, v in om.state_dict().items()})
    m = m.to(DEV)
    for static in (True, False):
        eng = GCNDeleteEngine(m, data.clone().to(DEV), neg.to(DEV), z_ori=zo.float().to(DEV), hoist_layer1=False,
                              static_negatives=static)
        got = eng.forward_backward().clone()
        want = torch.stack([loss.detach(), lr.detach(), ll.detach()])
        U.assert_close(got, want, what=f'losses (static_negatives={static})')
        U.assert_close(eng.params[0].grad, om.deletion1.deletion_weight.grad, what='dW_del1')
        U.assert_close(eng.params[1].grad, om.deletion2.deletion_weight.grad, what='dW_del2')
        again = eng.forward_backward().clone()
        assert torch.equal(again, got), 'an epoch is bitwise reproducible (no float atomics)'
