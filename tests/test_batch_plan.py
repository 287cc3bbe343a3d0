"""Host logic of the batched aggregation (gnndelete_b200/graph.py::BatchPlan), no GPU: a plain-Python
walk of the plan in exactly the order the kernel's workers follow it (csrc/spmm_batched.cu) must
reproduce the dense product, for any worker count — including counts that cut almost every row into
pieces — and for empty rows."""
import pytest
import torch

from gnndelete_b200.graph import BatchPlan


def _random_csr(n, avg, seed, hubs=2, empty_every=7):
    g = torch.Generator().manual_seed(seed)
    deg = torch.poisson(torch.full((n,), float(avg)), generator=g).long()
    deg[::empty_every] = 0
    for h in range(hubs):
        deg[(h * 37 + 5) % n] = 20 * avg + 13 * h
    rowptr = torch.cat([torch.zeros(1, dtype=torch.long), deg.cumsum(0)])
    nnz = int(rowptr[-1])
    col = torch.randint(0, n, (nnz,), generator=g)
    return rowptr.to(torch.int32), col.to(torch.int32), nnz


def _walk(plan, x, valp=None):
    """Emulates spmm_batched_kernel: every worker walks its contiguous batch range; flushes write a
    row or a piece; the pieces of a split row are added in piece order."""
    n, f = plan.num_rows, x.shape[1]
    out = torch.full((n, f), float('nan'), dtype=torch.float64)
    scratch = torch.zeros(max(plan.num_piece, 1), f, dtype=torch.float64)
    arrived = torch.zeros(max(plan.num_split, 1), dtype=torch.long)
    colp = plan.colp.view(-1, 8).long()
    desc = plan.desc.long()
    assert colp.shape[0] == plan.num_batches + 2 and desc.numel() == plan.num_batches + 2, 'two batches of slack'
    assert bool((colp[plan.num_batches:] == -1).all()) and bool((desc[plan.num_batches:] == 0).all())
    written = torch.zeros(n, dtype=torch.long)
    for w in range(plan.num_workers):
        b0 = w * plan.batches_per_worker
        b1 = min(b0 + plan.batches_per_worker, plan.num_batches)
        acc = torch.zeros(f, dtype=torch.float64)
        for b in range(b0, b1):
            for s in range(8):
                c = int(colp[b, s])
                if c >= 0:
                    wgt = 1.0 if valp is None else float(valp[b * 8 + s])
                    acc += wgt * x[c].double()
                else:
                    assert all(int(v) < 0 for v in colp[b, s:]), 'padding must be at the end of a batch'
            d = int(desc[b])
            if d < 0:
                ident = d & 0x3fffffff
                if d & 0x40000000:
                    scratch[ident] = acc
                    h = int(plan.piece_split[ident])
                    arrived[h] += 1
                    if arrived[h] == int(plan.split_npiece[h]):
                        p0 = int(plan.split_piece_beg[h])
                        row = int(plan.split_row[h])
                        out[row] = scratch[p0:p0 + int(plan.split_npiece[h])].sum(0)
                        written[row] += 1
                else:
                    out[ident] = acc
                    written[ident] += 1
                acc = torch.zeros(f, dtype=torch.float64)
        assert float(acc.abs().sum()) == 0.0, 'a worker range must end with a flush'
    assert bool((written == 1).all()), 'every row is written exactly once'
    return out


@pytest.mark.parametrize('workers', [1, 3, 16, 97, 100000])
def test_batch_plan_walk_matches_dense(workers):
    n = 150
    rowptr, col, nnz = _random_csr(n, 5, seed=workers)
    plan = BatchPlan(rowptr, col, n, nnz, workers)
    assert plan.num_workers * plan.batches_per_worker >= plan.num_batches
    x = torch.randn(n, 4, generator=torch.Generator().manual_seed(2))
    A = torch.zeros(n, n, dtype=torch.float64)
    rows = torch.repeat_interleave(torch.arange(n), (rowptr[1:] - rowptr[:-1]).long())
    A.index_put_((rows, col.long()), torch.ones(nnz, dtype=torch.float64), accumulate=True)
    torch.testing.assert_close(_walk(plan, x), A @ x.double(), rtol=1e-12, atol=1e-12)
    # weighted, values scattered through slot_of_entry (what the loss kernel does through pos_u / pos_v)
    val = torch.randn(nnz, generator=torch.Generator().manual_seed(3))
    valp = plan.pad_values(val)
    Aw = torch.zeros(n, n, dtype=torch.float64)
    Aw.index_put_((rows, col.long()), val.double(), accumulate=True)
    torch.testing.assert_close(_walk(plan, x, valp), Aw @ x.double(), rtol=1e-12, atol=1e-12)
    if workers >= 97:
        assert plan.num_split > 0, 'case must exercise split rows'


def test_batch_plan_slots_and_scale_weights():
    n = 64
    rowptr, col, nnz = _random_csr(n, 3, seed=11)
    plan = BatchPlan(rowptr, col, n, nnz, 10)
    soe = plan.slot_of_entry
    assert soe.numel() == nnz and soe.unique().numel() == nnz
    assert torch.equal(plan.colp[soe].long(), col.long())
    assert int((plan.colp >= 0).sum()) == nnz
    cs = torch.rand(n) + 0.5
    w = plan.col_scale_weights(cs)
    assert torch.equal(w[soe], cs[col.long()])
    assert w.numel() == plan.num_slots and float(w[plan.colp < 0].abs().sum()) == 0.0
