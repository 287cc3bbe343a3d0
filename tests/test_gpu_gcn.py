"""GPU parity: GCN path kernels and GCNDelete against the CPU oracle (fp64 truth)."""
import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def case(lib):
    return U.make_case('cora', 0.05)


def test_csr_build_bit_exact(lib, case):
    """CSR of the sdf edge set with PyG self-loop handling: bit-exact vs a CPU sort."""
    from gnndelete_b200.graph import build_csr
    shape, raw, df, data, neg = case
    ei = data.train_pos_edge_index[:, data.sdf_mask]
    # inject a few explicit self loops: they must be dropped and re-inserted once per node
    ei = torch.cat([ei, torch.tensor([[3, 5, 5], [3, 5, 5]])], 1)
    n = data.num_nodes
    csr = build_csr(ei[0].to(DEV), ei[1].to(DEV), n, self_loops=True)
    from oracle import pyg_ops as P
    full = P.add_remaining_self_loops(ei, n)
    key = full[1] * n + full[0]
    order = torch.sort(key, stable=True)[0]
    assert csr.nnz == full.shape[1]
    assert torch.equal(csr.col.cpu().long(), order % n)
    deg = torch.bincount(full[1], minlength=n)
    assert torch.equal(csr.rowptr.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), deg.cumsum(0)]))
    # eid maps back to the input column (or E + v for the inserted loop of node v)
    eid = csr.eid.cpu().long()
    real = eid < ei.shape[1]
    assert int((~real).sum()) == n
    assert torch.equal(eid[~real] - ei.shape[1], csr.col.cpu().long()[~real])
    assert torch.equal(ei[0][eid[real]], csr.col.cpu().long()[real])


def test_csr_rejects_out_of_range(lib):
    from gnndelete_b200.graph import build_csr
    ei = torch.tensor([[0, 1, 7], [1, 2, 0]], device=DEV)
    with pytest.raises(ValueError):
        build_csr(ei[0], ei[1], 4)


@pytest.mark.parametrize('feat', [128, 64, 32, 100, 7])
def test_spmm_matches_dense(lib, case, feat):
    from gnndelete_b200 import ops
    from gnndelete_b200.graph import build_csr
    shape, raw, df, data, neg = case
    ei = data.train_pos_edge_index
    n = data.num_nodes
    csr = build_csr(ei[0].to(DEV), ei[1].to(DEV), n, self_loops=False, seg_len=32)
    assert csr.num_seg > 0, 'case must exercise the long-row split path'
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, feat, generator=g)
    rs, cs = torch.rand(n, generator=g) + 0.5, torch.rand(n, generator=g) + 0.5
    bias = torch.randn(feat, generator=g)
    A = torch.zeros(n, n, dtype=torch.float64)
    A.index_put_((ei[1], ei[0]), torch.ones(ei.shape[1], dtype=torch.float64), accumulate=True)
    ref = rs.double().view(-1, 1) * (A @ (cs.double().view(-1, 1) * x.double())) + 0.5 * x.double() + bias.double()
    out = ops.spmm(csr, x.to(DEV), col_scale=cs.to(DEV), row_scale=rs.to(DEV), self_coef=0.5, bias=bias.to(DEV))
    U.assert_close(out, ref, what=f'spmm F={feat}')
    ref2 = A @ x.double()
    out2 = ops.spmm(csr, x.to(DEV))
    U.assert_close(out2, ref2, what=f'spmm unweighted F={feat}')


def test_spmm_two_pass_128_wide(lib, case, monkeypatch):
    """A 128-wide aggregation as two 64-wide column passes (`ops.SPLIT128`, on by default for sources about the size of
    the L2, forced here at test size): same result as the dense product and as the single pass, weighted (with self term and bias) and unweighted."""
    from gnndelete_b200 import ops
    from gnndelete_b200.graph import build_csr
    shape, raw, df, data, neg = case
    ei = data.train_pos_edge_index
    n = data.num_nodes
    csr = build_csr(ei[0].to(DEV), ei[1].to(DEV), n, self_loops=False)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(n, 128, generator=g)
    rs, cs = torch.rand(n, generator=g) + 0.5, torch.rand(n, generator=g) + 0.5
    bias = torch.randn(128, generator=g)
    A = torch.zeros(n, n, dtype=torch.float64)
    A.index_put_((ei[1], ei[0]), torch.ones(ei.shape[1], dtype=torch.float64), accumulate=True)
    ref_w = rs.double().view(-1, 1) * (A @ (cs.double().view(-1, 1) * x.double())) + 0.5 * x.double() + bias.double()
    ref_u = A @ x.double()
    xd, rsd, csd, bd = x.to(DEV), rs.to(DEV), cs.to(DEV), bias.to(DEV)

    def run():
        w = ops.spmm(csr, xd, col_scale=csd, row_scale=rsd, self_coef=0.5, bias=bd)
        u = ops.spmm(csr, xd)
        return w, u

    monkeypatch.setattr(ops, 'SPLIT128', True)
    monkeypatch.setattr(ops, 'SPLIT128_SOURCE_BYTES', (0, float('inf')))
    run()                                                       # builds the 64-wide plans / cached weights
    c0 = lib.gd_launch_count()
    w2, u2 = run()
    n_two = lib.gd_launch_count() - c0
    monkeypatch.setattr(ops, 'SPLIT128', False)
    run()
    c0 = lib.gd_launch_count()
    w1, u1 = run()
    n_one = lib.gd_launch_count() - c0
    assert n_two > n_one, 'the two-pass path was not taken'
    U.assert_close(w2, ref_w, what='two-pass weighted')
    U.assert_close(u2, ref_u, what='two-pass unweighted')
    # same sums in a different piece order (the 64- and 128-wide plans cut long rows at different places)
    U.assert_close(w2, w1, tol=5e-6, what='two-pass vs single pass')
    U.assert_close(u2, u1, tol=5e-6, what='two-pass vs single pass, unweighted')


@pytest.mark.parametrize('workers', [1, 7, 1000, 9472, 10 ** 6])
def test_batch_plan_native_builder_bit_exact(lib, case, workers):
    """The CUDA plan builder and the tensor-op builder (whose walk tests/test_batch_plan.py checks on the CPU)
    produce identical arrays, for worker counts from 1 to more than there are batches."""
    from gnndelete_b200.graph import BatchPlan, build_csr
    shape, raw, df, data, neg = case
    n = data.num_nodes
    ei = data.train_pos_edge_index
    ei = ei[:, ei[1] % 5 != 0]                       # empty rows
    csr = build_csr(ei[0].to(DEV), ei[1].to(DEV), n, self_loops=False)
    a = BatchPlan(csr.rowptr, csr.col, n, csr.nnz, workers, native=True)
    b = BatchPlan(csr.rowptr, csr.col, n, csr.nnz, workers, native=False)
    for k in ('num_batches', 'num_workers', 'batches_per_worker', 'num_slots', 'num_split', 'num_piece'):
        assert getattr(a, k) == getattr(b, k), k
    for k in ('desc', 'colp', 'slot_of_entry', 'piece_split', 'split_row', 'split_piece_beg', 'split_npiece'):
        x, y = getattr(a, k), getattr(b, k)
        assert (x is None) == (y is None), k
        if x is not None:
            assert torch.equal(x.long(), y.long()), k


@pytest.mark.parametrize('feat', [128, 64, 32])
@pytest.mark.parametrize('workers', [7, 1000, 0])
def test_spmm_batched_matches_dense(lib, case, feat, workers):
    """gd_spmm_batched on plans balanced for few / very many (rows cut into pieces) / the resident
    number of workers; unweighted, padded weights, accumulate, empty rows."""
    from gnndelete_b200 import _lib as L
    from gnndelete_b200.graph import BatchPlan, build_csr
    shape, raw, df, data, neg = case
    n = data.num_nodes
    ei = data.train_pos_edge_index
    ei = ei[:, ei[1] % 5 != 0]                       # rows 0, 5, 10, ... are empty
    csr = build_csr(ei[0].to(DEV), ei[1].to(DEV), n, self_loops=False)
    if workers == 0:
        workers = lib.gd_spmm_batched_workers(feat, 1)
        assert workers >= 148 * 8
    bp = BatchPlan(csr.rowptr, csr.col, n, csr.nnz, workers)
    if workers == 1000:
        assert bp.num_split > 0
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, feat, generator=g)
    rs = torch.rand(n, generator=g) + 0.5
    bias = torch.randn(feat, generator=g)
    val = torch.randn(csr.nnz, generator=g)
    A = torch.zeros(n, n, dtype=torch.float64)
    A.index_put_((ei[1], ei[0]), torch.ones(ei.shape[1], dtype=torch.float64), accumulate=True)
    xd = x.to(DEV)

    def run(valp, row_scale, self_coef, b, out, acc):
        L.call('gd_spmm_batched', bp.ref, L.ptr(valp), L.ptr(row_scale), L.ptr(xd), xd.stride(0), feat, self_coef,
               L.ptr(b), L.ptr(out), out.stride(0), L.ptr(bp.scratch(feat)), acc, L.stream())
        return out

    out = run(None, rs.to(DEV), 0.5, bias.to(DEV), torch.empty(n, feat, device=DEV), 0)
    ref = rs.double().view(-1, 1) * (A @ x.double()) + 0.5 * x.double() + bias.double()
    U.assert_close(out, ref, what=f'batched F={feat} W={workers}')
    assert torch.equal(out.cpu()[0], (0.5 * x[0] + bias)), 'empty row = self term + bias'
    # per-entry weights in CSR order -> padded layout; twice into the same buffer = accumulate
    src, dst = csr.col.cpu().long(), torch.repeat_interleave(torch.arange(n), (csr.rowptr[1:] - csr.rowptr[:-1]).cpu().long())
    Aw = torch.zeros(n, n, dtype=torch.float64)
    Aw.index_put_((dst, src), val.double(), accumulate=True)
    valp = bp.pad_values(val.to(DEV))
    out2 = run(valp, None, 0.0, None, torch.empty(n, feat, device=DEV), 0)
    U.assert_close(out2, Aw @ x.double(), what=f'batched weighted F={feat} W={workers}')
    again = run(valp, None, 0.0, None, out2.clone(), 1)
    U.assert_close(again, 2 * (Aw @ x.double()), what='batched accumulate')
    # a second plain CSR added row by row at the flush (this step's negative pairs in the loss gradient)
    tm = 3 * n
    ts_, td_ = torch.randint(0, n, (tm,), generator=g), torch.randint(0, n, (tm,), generator=g)
    tail = build_csr(ts_.to(DEV), td_.to(DEV), n, self_loops=False)
    tval = torch.randn(tm, generator=g)
    tval_csr = tval.to(DEV)[tail.eid.long()]
    T = torch.zeros(n, n, dtype=torch.float64)
    T.index_put_((td_, ts_), tval.double(), accumulate=True)
    out3 = torch.empty(n, feat, device=DEV)
    L.call('gd_spmm_batched_tail', bp.ref, L.ptr(valp), L.ptr(tail.rowptr), L.ptr(tail.col), L.ptr(tval_csr), None, L.ptr(xd),
           xd.stride(0), feat, 0.0, None, L.ptr(out3), out3.stride(0), L.ptr(bp.scratch(feat)), 0, L.stream())
    U.assert_close(out3, (Aw + T) @ x.double(), what=f'batched + tail CSR F={feat} W={workers}')
    # bitwise reproducible (piece order is fixed by the plan, not by arrival)
    assert torch.equal(run(valp, None, 0.0, None, torch.empty(n, feat, device=DEV), 0), out2)


@pytest.mark.parametrize('m,k,n,nk', [(300, 128, 128, True), (1000, 128, 64, True), (257, 500, 128, True),
                                      (513, 64, 64, False), (100, 30, 17, False)])
def test_gemm_rows(lib, m, k, n, nk):
    from gnndelete_b200 import ops
    g = torch.Generator().manual_seed(2)
    a = torch.randn(m, k, generator=g)
    b = torch.randn((n, k) if nk else (k, n), generator=g)
    bias = torch.randn(n, generator=g)
    s = torch.rand(m, generator=g) + 0.5
    bm = b.double().t() if nk else b.double()
    ref = (a.double().clamp(min=0) @ bm + bias.double()) * s.double().view(-1, 1)
    out = ops.gemm_rows(a.to(DEV), b.to(DEV), nk, bias=bias.to(DEV), out_scale=s.to(DEV), relu_in=True)
    U.assert_close(out, ref, what='gemm_rows')
    # gathered rows, in place
    rows = torch.randperm(m, generator=g)[: m // 3].sort()[0]
    out2 = torch.full((m, n), -7.0, device=DEV)
    ops.gemm_rows(a.to(DEV), b.to(DEV), nk, out=out2, rows=rows.to(DEV).int())
    ref2 = torch.full((m, n), -7.0, dtype=torch.float64)
    ref2[rows] = a.double()[rows] @ bm
    U.assert_close(out2, ref2, what='gemm_rows gathered')


@pytest.mark.parametrize('kernel', ['wt', 'ring'])
@pytest.mark.parametrize('m,k,n,nk', [(40001, 128, 128, True), (30000, 128, 64, True), (20011, 64, 64, False), (25000, 64, 128, False),
                                      (130, 32, 96, True), (5000, 96, 64, True), (63, 128, 128, False)])
def test_gemm_rows_epoch_epilogues(lib, m, k, n, nk, kernel, monkeypatch):
    """The epilogues of the Del-training epoch on both tcgen05 row GEMMs (weights in tensor memory, gemm_tc_wt.cu, and
    the shared-memory ring, gemm_tc.cu): row scale, ReLU prologue, bit-packed ReLU mask out, gate bits in, gathered rows;
    sizes with several tiles per CTA (ring wrap-around, both accumulator buffers) and ragged last tiles."""
    from gnndelete_b200 import ops
    monkeypatch.setenv('GD_GEMM_ROWS', kernel)
    g = torch.Generator().manual_seed(m + k + n)
    a = torch.randn(m, k, generator=g)
    b = torch.randn((n, k) if nk else (k, n), generator=g)
    s = torch.rand(m, generator=g) + 0.5
    bm = b.double().t() if nk else b.double()
    ad, bd, sd = a.to(DEV), b.to(DEV), s.to(DEV)
    # (1) all rows: ReLU prologue + row scale
    out = ops.gemm_rows(ad, bd, nk, out_scale=sd, relu_in=True)
    U.assert_close(out, (a.double().clamp(min=0) @ bm) * s.double().view(-1, 1), what='scale + relu_in')
    # (2) gathered rows in place + the ReLU mask of the result as bits
    rows = torch.randperm(m, generator=g)[: (2 * m) // 3].sort()[0]
    bits = torch.full((m, n // 32), -1, dtype=torch.int32, device=DEV)
    out2 = torch.full((m, n), -7.0, device=DEV)
    ops.gemm_rows(ad, bd, nk, out=out2, rows=rows.to(DEV).int(), relu_mask_out=bits)
    ref2 = torch.full((m, n), -7.0, dtype=torch.float64)
    prod = a.double()[rows] @ bm
    ref2[rows] = prod
    U.assert_close(out2, ref2, what='gathered rows')
    got = ((bits[rows.to(DEV)].unsqueeze(-1) >> torch.arange(32, device=DEV)) & 1).reshape(rows.numel(), n).bool().cpu()
    sure = prod.abs() > 1e-4 * prod.abs().max()                      # sign of near-zero products may differ in the last ulp
    assert torch.equal(got[sure], (prod > 0)[sure]), 'ReLU mask bits'
    assert torch.equal(bits[~torch.isin(torch.arange(m), rows).to(DEV)], torch.full((m - rows.numel(), n // 32), -1, dtype=torch.int32, device=DEV))
    # (3) gate bits in + row scale (the dX1 GEMM): out = scale * (a . B) where the gate bit is set, else 0
    gate = torch.rand(m, n, generator=g) > 0.4
    w = (gate.view(m, n // 32, 32).long() << torch.arange(32)).sum(-1)
    gbits = torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32).to(DEV)
    out3 = torch.full((m, n), 3.0, device=DEV)
    ops.gemm_rows(ad, bd, nk, out=out3, rows=rows.to(DEV).int(), out_scale=sd, gate_bits=gbits)
    ref3 = torch.full((m, n), 3.0, dtype=torch.float64)
    ref3[rows] = torch.where(gate[rows], prod * s.double()[rows].view(-1, 1), torch.zeros_like(prod))
    U.assert_close(out3, ref3, what='gate bits + scale')
    # ... and the ReLU mask of a gated result: bit = gate bit AND (value > 0)
    bits3 = torch.zeros(m, n // 32, dtype=torch.int32, device=DEV)
    ops.gemm_rows(ad, bd, nk, out=out3, rows=rows.to(DEV).int(), gate_bits=gbits, relu_mask_out=bits3)
    got3 = ((bits3[rows.to(DEV)].unsqueeze(-1) >> torch.arange(32, device=DEV)) & 1).reshape(rows.numel(), n).bool().cpu()
    assert torch.equal(got3[sure], (gate[rows] & (prod > 0))[sure]), 'mask bits of a gated result'
    # (4) mask bits with a row scale (sign taken after scaling, scale > 0)
    bits4 = torch.zeros(m, n // 32, dtype=torch.int32, device=DEV)
    out4 = ops.gemm_rows(ad, bd, nk, out_scale=sd, relu_mask_out=bits4)
    full = a.double() @ bm
    U.assert_close(out4, full * s.double().view(-1, 1), what='scale + mask bits')
    got4 = ((bits4.unsqueeze(-1) >> torch.arange(32, device=DEV)) & 1).reshape(m, n).bool().cpu()
    sure4 = full.abs() > 1e-4 * full.abs().max()
    assert torch.equal(got4[sure4], (full > 0)[sure4])
    # (5) bias + ReLU epilogue (the conv layers of GIN / original-model training), all rows and gathered rows
    bias = torch.randn(n, generator=g)
    out5 = ops.gemm_rows(ad, bd, nk, bias=bias.to(DEV), relu_out=True)
    U.assert_close(out5, (full + bias.double()).clamp(min=0), what='bias + relu_out')
    out6 = torch.full((m, n), 2.0, device=DEV)
    ops.gemm_rows(ad, bd, nk, out=out6, rows=rows.to(DEV).int(), bias=bias.to(DEV))
    ref6 = torch.full((m, n), 2.0, dtype=torch.float64)
    ref6[rows] = prod + bias.double()
    U.assert_close(out6, ref6, what='gathered rows + bias')


@pytest.mark.parametrize('m,k1,n2', [(5000, 128, 128), (777, 64, 64), (300, 100, 30)])
def test_gemm_tn_rows(lib, m, k1, n2):
    from gnndelete_b200 import ops
    g = torch.Generator().manual_seed(3)
    a, gr = torch.randn(m, k1, generator=g), torch.randn(m, n2, generator=g)
    rows = torch.randperm(m, generator=g)[: m // 2].sort()[0]
    ref = a.double()[rows].t() @ gr.double()[rows]
    out = ops.gemm_tn_rows(a.to(DEV), gr.to(DEV), rows=rows.to(DEV).int())
    U.assert_close(out, ref, what='gemm_tn_rows')


@pytest.mark.parametrize('kernel', ['wt', 'ring'])
@pytest.mark.parametrize('m,k1,n2', [(60001, 128, 128), (50000, 64, 64), (45000, 128, 64), (45000, 64, 128), (9000, 96, 32), (100, 128, 128)])
def test_gemm_tn_rows_both_tensor_core_kernels(lib, m, k1, n2, kernel, monkeypatch):
    """The weight-gradient contraction on both tcgen05 kernels (A^T in tensor memory, gemm_tn_wt.cu, and both operands in
    shared memory, gemm_tc.cu): gathered rows, all rows, ReLU prologue and row scale; sizes with several flush groups per
    CTA, ragged last stages and a single partial stage."""
    from gnndelete_b200 import ops
    monkeypatch.setenv('GD_GEMM_TN', kernel)
    g = torch.Generator().manual_seed(m + k1 + n2)
    a, gr = torch.randn(m, k1, generator=g), torch.randn(m, n2, generator=g)
    sc = torch.rand(m, generator=g) + 0.5
    rows = torch.randperm(m, generator=g)[: (2 * m) // 3].sort()[0]
    ad, gd_, rd, sd = a.to(DEV), gr.to(DEV), rows.to(DEV).int(), sc.to(DEV)
    U.assert_close(ops.gemm_tn_rows(ad, gd_, rows=rd), a.double()[rows].t() @ gr.double()[rows], what='gathered rows')
    U.assert_close(ops.gemm_tn_rows(ad, gd_), a.double().t() @ gr.double(), what='all rows')
    ref = (a.double()[rows].clamp(min=0) * sc.double()[rows].view(-1, 1)).t() @ gr.double()[rows]
    U.assert_close(ops.gemm_tn_rows(ad, gd_, rows=rd, relu_a=True, a_scale=sd), ref, what='relu + row scale')
    # deterministic: partials are reduced in CTA order
    assert torch.equal(ops.gemm_tn_rows(ad, gd_, rows=rd), ops.gemm_tn_rows(ad, gd_, rows=rd))


@pytest.mark.parametrize('m,k,n,k1', [(60001, 64, 128, 128), (130000, 64, 128, 128), (9000, 32, 64, 96), (20000, 64, 96, 64),
                                      (300, 64, 128, 128), (64, 32, 32, 32)])
def test_gemm_dxdw_chained_kernel(lib, m, k, n, k1):
    """Input gradient chained into the weight gradient through tensor memory (gemm_dxdw_wt.cu) vs fp64 and vs the two-kernel
    path it replaces: gathered rows with row scale and gate bits, all rows without, several flush groups per CTA, ragged
    last tiles, a single partial tile."""
    from gnndelete_b200 import ops
    g = torch.Generator().manual_seed(m + k + n + k1)
    x, a = torch.randn(m, k, generator=g), torch.randn(m, k1, generator=g)
    w = torch.randn(k, n, generator=g) / k ** 0.5
    sc = torch.rand(m, generator=g) + 0.5
    gate = torch.rand(m, n, generator=g) > 0.4
    rows = torch.randperm(m, generator=g)[: (2 * m) // 3].sort()[0]
    bits = torch.zeros(m, n // 32, dtype=torch.int64)
    for j in range(n):
        bits[:, j // 32] |= gate[:, j].long() << (j % 32)
    bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).int()
    xd, ad, wd, sd, bd, rd = x.to(DEV), a.to(DEV), w.to(DEV), sc.to(DEV), bits.to(DEV), rows.to(DEV).int()
    assert ops.gemm_dxdw_supported(xd, wd, False, ad)
    dx = ((x.double() * sc.double().view(-1, 1)) @ w.double()) * gate.double()
    ref = a.double()[rows].t() @ dx[rows]
    out = ops.gemm_dxdw(xd, wd, False, ad, rows=rd, in_scale=sd, gate_bits=bd)
    U.assert_close(out, ref, what='gathered rows, scale, gate bits')
    dxd = torch.zeros(m, n, device=DEV)
    ops.gemm_rows(xd, wd, False, out=dxd, rows=rd, out_scale=sd, gate_bits=bd)
    two = ops.gemm_tn_rows(ad, dxd, rows=rd)
    assert (out - two).abs().max().item() <= 2e-5 * ref.abs().max().item()
    U.assert_close(ops.gemm_dxdw(xd, wd.t().contiguous(), True, ad), a.double().t() @ (x.double() @ w.double()), what='all rows, B as [n, k]')
    assert torch.equal(ops.gemm_dxdw(xd, wd, False, ad, rows=rd, in_scale=sd, gate_bits=bd), out)      # deterministic


def _load(gnn, om, shape, data, **kw):
    from gnndelete_b200 import models as M
    cls = {'gcn': M.GCNDelete, 'gin': M.GINDelete}[gnn]
    m = cls(U.args_for(shape), data.sdf_node_1hop_mask, data.sdf_node_2hop_mask, **kw)
    missing = m.load_state_dict({k: v.float() for k, v in om.state_dict().items()}, strict=True)
    return m.to(DEV)


@pytest.mark.parametrize('gnn', ['gcn', 'gin'])
def test_delete_model_forward_and_grads(lib, case, gnn):
    """z1, z2, decode logits, edge-form loss and deletion_weight.grad vs the oracle."""
    from oracle import unlearn as OU
    shape, raw, df, data, neg = case
    om = U.oracle_model(gnn, shape, data, dtype=torch.float64)
    d64 = data.clone()
    d64.x = data.x.double()
    with torch.no_grad():
        zo = om.get_original_embeddings(d64.x, d64.train_pos_edge_index[:, d64.dr_mask])
    loss_o, lr_o, ll_o, z_o = OU.edge_form_loss(om, d64, neg, zo)
    loss_o.backward()

    m = _load(gnn, om, shape, data)
    dd = data.clone().to(DEV)
    ei = dd.train_pos_edge_index[:, dd.sdf_mask]
    z1, z2 = m(dd.x, ei, return_all_emb=True)
    z1_o, z2_o = om(d64.x, d64.train_pos_edge_index[:, d64.sdf_mask], return_all_emb=True)
    U.assert_close(z1, z1_o, what='z1')
    U.assert_close(z2, z2_o, what='z2')
    # original embeddings on the dr edge set
    zo_g = m.get_original_embeddings(dd.x, dd.train_pos_edge_index[:, dd.dr_mask])
    U.assert_close(zo_g, zo, what='z_ori')
    # decode + losses through the reference-shaped autograd path
    negd = neg.to(DEV)
    logits = m.decode(z2, dd.train_pos_edge_index[:, dd.df_mask], negd)
    U.assert_close(logits, om.decode(z2_o, d64.train_pos_edge_index[:, d64.df_mask], neg), what='df_logits')
    n = int(dd.df_mask.sum())
    loss_r = torch.nn.functional.mse_loss(logits[:n], logits[n:])
    edge = ei
    lower = edge[0] < edge[1]
    row, col = edge[0][lower], edge[1][lower]
    lg = m.decode(z2, torch.stack([row, col]))
    lo = m.decode(zo_g, torch.stack([row, col])).detach()
    loss_l = torch.nn.functional.mse_loss(lg, lo)
    loss = 0.5 * loss_r + 0.5 * loss_l
    U.assert_close(loss_r, lr_o, what='loss_r')
    U.assert_close(loss_l, ll_o, what='loss_l')
    loss.backward()
    U.assert_close(m.deletion2.deletion_weight.grad, om.deletion2.deletion_weight.grad, what='dW_del2')
    U.assert_close(m.deletion1.deletion_weight.grad, om.deletion1.deletion_weight.grad, what='dW_del1')


@pytest.mark.parametrize('dim,static,det', [(64, False, False), (64, False, True), (64, True, False), (128, False, False),
                                            (128, False, True), (32, True, False), (48, False, False)])
def test_fused_edge_loss(lib, case, dim, static, det):
    """Fused decoder + DEC / NI losses + dz == autograd of the oracle's loss w.r.t. z: the node-side one-pass kernel
    (gd_node_loss_fwd_bwd, widths 32 / 64 / 128) and the pair-side kernel + incidence gather (any other width), with
    the negatives in the fixed incidence (static), in the per-step tail CSR (deterministic) or added with vector float
    reductions (the default for replaceable negatives)."""
    from gnndelete_b200.losses import EdgeLossPlan
    shape, raw, df, data, neg = case
    g = torch.Generator().manual_seed(5)
    n = data.num_nodes
    z = torch.randn(n, dim, generator=g, dtype=torch.float64, requires_grad=True)
    zo = torch.randn(n, dim, generator=g, dtype=torch.float64)
    ei = data.train_pos_edge_index
    dfe = ei[:, data.df_mask]
    sdf = ei[:, data.sdf_mask]
    ni = sdf[:, sdf[0] < sdf[1]]
    nd = dfe.shape[1]
    lg = (z[torch.cat([dfe[0], neg[0]])] * z[torch.cat([dfe[1], neg[1]])]).sum(-1)
    loss_r = torch.nn.functional.mse_loss(lg[:nd], lg[nd:])
    loss_l = torch.nn.functional.mse_loss((z[ni[0]] * z[ni[1]]).sum(-1), (zo[ni[0]] * zo[ni[1]]).sum(-1))
    loss = 0.5 * loss_r + 0.5 * loss_l
    loss.backward()
    zg = z.detach().float().to(DEV)
    plan = EdgeLossPlan(dfe.to(DEV), neg.to(DEV), ni.to(DEV), n, z_ori=zo.float().to(DEV), static_negatives=static,
                        deterministic=det)
    assert plan.mode == ('node' if dim in (32, 64, 128) else 'pair')
    assert plan.neg_atomic == (plan.mode == 'node' and not static and not det)
    losses = plan.forward(zg).clone()
    dz = plan.backward(zg).clone()
    U.assert_close(losses, torch.stack([loss, loss_r, loss_l]), what='losses')
    U.assert_close(dz, z.grad, what='dz')
    U.assert_close(plan.logits[:2 * nd], lg, what='logits')
    # deterministic: bitwise identical on repeat
    l2 = plan.forward(zg).clone()
    dz2 = plan.backward(zg)
    assert torch.equal(l2, losses)
    if plan.neg_atomic:          # float reductions: the order of a row's few negative contributions is not fixed
        U.assert_close(dz2, dz, tol=1e-6, what='dz on repeat')
    else:
        assert torch.equal(dz, dz2)
    # replaced negatives (gnndelete.py:221-225): same check against the oracle on the new set
    neg2 = torch.randint(0, n, neg.shape, generator=g)
    z.grad = None
    lg2 = (z[torch.cat([dfe[0], neg2[0]])] * z[torch.cat([dfe[1], neg2[1]])]).sum(-1)
    loss_r2 = torch.nn.functional.mse_loss(lg2[:nd], lg2[nd:])
    loss_l2 = torch.nn.functional.mse_loss((z[ni[0]] * z[ni[1]]).sum(-1), (zo[ni[0]] * zo[ni[1]]).sum(-1))
    (0.5 * loss_r2 + 0.5 * loss_l2).backward()
    plan.update_negatives(neg2.to(DEV))
    plan.check_negatives()
    losses2 = plan.forward(zg).clone()
    U.assert_close(losses2, torch.stack([0.5 * loss_r2 + 0.5 * loss_l2, loss_r2, loss_l2]).detach(), what='losses (new negatives)')
    U.assert_close(plan.backward(zg), z.grad, what='dz (new negatives)')


def test_deletion_layer_semantics(lib):
    """mask=None -> stored mask; both None -> identity; returns a new tensor; index-tensor mask."""
    from gnndelete_b200.models import DeletionLayer
    g = torch.Generator().manual_seed(6)
    x = torch.randn(50, 64, generator=g).to(DEV)
    mask = torch.rand(50, generator=g) < 0.4
    lay = DeletionLayer(64, mask).to(DEV)
    y = lay(x)
    w = lay.deletion_weight.detach()
    ref = x.clone()
    ref[mask.to(DEV)] = x[mask.to(DEV)] @ w
    U.assert_close(y, ref, what='stored mask')
    assert y.data_ptr() != x.data_ptr()
    assert DeletionLayer(64, None).to(DEV)(x) is x
    other = torch.rand(50, generator=g) < 0.5
    y2 = lay(x, other.to(DEV))
    ref2 = x.clone()
    ref2[other.to(DEV)] = x[other.to(DEV)] @ w
    U.assert_close(y2, ref2, what='call mask')
    idx = other.nonzero().squeeze(1).to(DEV)
    U.assert_close(lay(x, idx), ref2, what='index mask')


def test_adam_matches_torch(lib):
    from gnndelete_b200 import ops
    g = torch.Generator().manual_seed(7)
    p0 = torch.randn(64, 64, generator=g)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=1e-3)
    p = p0.clone().to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, device=DEV)
    for i in range(5):
        gr = torch.randn(64, 64, generator=g)
        p_ref.grad = gr.clone()
        opt.step()
        ops.adam_step(p, gr.to(DEV), m, v, step, 1e-3)
    assert step.item() == 5
    U.assert_close(p, p_ref.detach(), tol=1e-6, what='adam')
