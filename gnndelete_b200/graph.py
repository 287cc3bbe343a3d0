"""Device-side graph structures behind the conv layers: destination-major CSR, its
transpose for the backward pass, GCN normalisation and the long-row split plan.

The reference rebuilds the equivalent state inside every PyG conv call
(``gcn_norm`` with ``cached=False``, ``add_remaining_self_loops``; gcn.py:11-12).
The edge set of an unlearning run is fixed (``train_pos_edge_index[:, sdf_mask]``,
gnndelete.py:215), so the structures are built once by the CUDA builders in
``csrc/graph_build.cu`` and cached.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib as L

SEG_LEN = 128    # rows longer than this are split into segments (spmm.cu)
DEG_SORT_WINDOW = 4096
# GD_SPMM=pipe selects the older row-walking kernel (A/B measurements, tests/test_gpu_fullsize.py); default: batched
BATCHED = os.environ.get('GD_SPMM', 'batched')[0] == 'b'
OVERSUB = int(os.environ.get('GD_SPMM_OVERSUB', '1'))


class CSR:
    """Destination-major CSR on the device + the ``gd_csr_t`` the kernels take."""

    def __init__(self, rowptr, col, eid, rel, num_rows, nnz, plan=None, row_perm=None, grp_row=None):
        self.rowptr, self.col, self.eid, self.rel = rowptr, col, eid, rel
        self.row_perm = row_perm
        self.grp_row = grp_row
        self.num_rows, self.nnz = int(num_rows), int(nnz)
        self.plan = plan or {}
        self._scratch = {}
        self._bplans = {}
        self.dynamic = False     # True: the arrays are rewritten in place (no cached batch plan)
        # batch-plan workers per resident sub-warp.  1: one persistent wave (fastest alone); > 1: more, shorter CTAs, for
        # kernels that share the GPU with a collective or another kernel (a persistent wave that does not fit at once
        # runs its left-over CTAs as a second wave of the same length)
        self.oversub = OVERSUB
        s = L.CsrStruct()
        s.num_rows, s.nnz = self.num_rows, self.nnz
        s.rowptr, s.col = rowptr.data_ptr(), col.data_ptr()
        if plan:
            s.seg_len, s.num_heavy, s.num_seg = plan['seg_len'], plan['num_heavy'], plan['num_seg']
            for k in ('heavy_row', 'heavy_seg_beg', 'heavy_nseg', 'seg_row', 'seg_beg', 'seg_heavy', 'heavy_ticket'):
                setattr(s, k, plan[k].data_ptr())
        if row_perm is not None:
            s.row_perm = row_perm.data_ptr()
        if grp_row is not None and grp_row.numel() > 1:
            s.grp_row = grp_row.data_ptr()
            s.num_grp = grp_row.numel() - 1
        self.struct = s
        self.ref = C.byref(s)

    @property
    def num_seg(self):
        return self.plan.get('num_seg', 0)

    def bplan(self, feat, weighted=False, bf16=False):
        """Batch plan balanced for the sub-warps resident at this width (built on first use);
        ``None`` when the batched kernel does not cover the width.  ``bf16``: the plan of the bf16-source kernel
        (``gd_spmm_batched_bf16``: feat / 8 lanes per sub-warp, widths 64 / 128)."""
        if feat not in ((64, 128) if bf16 else (32, 64, 128)) or self.num_rows == 0 or self.num_rows >= (1 << 30) \
                or not BATCHED or self.dynamic or self.nnz == 0:
            return None
        if bf16:
            workers = L.load().gd_spmm_batched_bf16_workers(int(feat), int(bool(weighted))) * self.oversub
        else:
            workers = L.load().gd_spmm_batched_workers(int(feat), int(bool(weighted))) * self.oversub
        bp = self._bplans.get(workers)
        if bp is None:
            bp = BatchPlan(self.rowptr, self.col, self.num_rows, self.nnz, workers)
            self._bplans[workers] = bp
        return bp

    def gat_scratch(self, channels):
        """Scratch of the GAT kernels' long-row segments (``gd_gat_scratch_floats``); ``None`` without a split plan."""
        if self.num_seg == 0:
            return None
        key = ('gat', channels)
        buf = self._scratch.get(key)
        if buf is None:
            buf = torch.empty(self.num_seg * (channels + 4), dtype=torch.float32, device=self.rowptr.device)   # 16-byte aligned records
            self._scratch[key] = buf
        return buf

    def scratch(self, feat):
        """Per-width scratch for the split-row partial sums (allocated once)."""
        if self.num_seg == 0:
            return None
        buf = self._scratch.get(feat)
        if buf is None:
            buf = torch.empty(self.num_seg * feat, dtype=torch.float32, device=self.rowptr.device)
            self._scratch[feat] = buf
        return buf


class BatchPlan:
    """Batch plan of a CSR for ``gd_spmm_batched`` (csrc/spmm_batched.cu): rows cut into batches of
    8 padded column slots, the batch list cut into ``num_workers`` equal contiguous ranges, rows
    straddling a range boundary cut into pieces.  Built once per edge set with tensor ops on the
    device (plan time, not on the epoch path).

    ``slot_of_entry[k]`` is the padded slot of CSR entry ``k``: callers that produce per-entry
    values (the loss backward) write them straight into the padded layout."""

    SLOTS = 8
    FLUSH, PIECE = -(1 << 31), 1 << 30

    def __init__(self, rowptr, col, num_rows, nnz, num_workers, native=None):
        """``native``: build with the CUDA builder (``gd_spmm_bplan_count`` / ``_fill``; default on CUDA tensors) or with
        the tensor-op builder below (the CPU-testable statement of the layout; both produce identical arrays)."""
        self.num_rows = int(num_rows)
        if native is None:
            native = rowptr.is_cuda and self.num_rows > 0
        if native:
            self._build_native(rowptr, col, int(nnz), int(num_workers))
        else:
            self._build_torch(rowptr, col, int(nnz), int(num_workers))
        self._finish()

    def _build_native(self, rowptr, col, nnz, num_workers):
        dev, n, S = rowptr.device, self.num_rows, self.SLOTS
        lib = L.load()
        ws_bytes = lib.gd_spmm_bplan_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        sizes = torch.zeros(3, dtype=torch.int32, device=dev)
        L.call('gd_spmm_bplan_count', L.ptr(rowptr, 'i32'), n, int(num_workers), L.ptr(sizes), L.ptr(ws), ws_bytes, L.stream())
        nb, nsplit, npiece = sizes.tolist()
        self.num_batches = nb
        self.num_workers = max(1, min(int(num_workers), max(nb, 1)))
        per = -(-nb // self.num_workers) if nb else 1
        self.batches_per_worker = per
        self.num_workers = -(-nb // per) if nb else 1
        self.num_slots = (nb + 2) * S
        i32 = dict(dtype=torch.int32, device=dev)
        self.desc = torch.empty(nb + 2, **i32)
        self.colp = torch.empty(self.num_slots, **i32)
        self.slot_of_entry = torch.empty(max(nnz, 1), dtype=torch.int64, device=dev)
        self.num_split, self.num_piece = nsplit, npiece
        if nsplit:
            self.piece_split = torch.empty(npiece, **i32)
            self.split_row = torch.empty(nsplit, **i32)
            self.split_piece_beg = torch.empty(nsplit, **i32)
            self.split_npiece = torch.empty(nsplit, **i32)
            self.split_ticket = torch.zeros(nsplit, **i32)
        else:
            self.piece_split = self.split_row = self.split_npiece = self.split_piece_beg = self.split_ticket = None
        L.call('gd_spmm_bplan_fill', L.ptr(rowptr, 'i32'), L.ptr(col, 'i32'), n, nb, per, L.ptr(self.desc), L.ptr(self.colp),
               L.ptr(self.slot_of_entry), L.ptr(self.piece_split), L.ptr(self.split_row), L.ptr(self.split_piece_beg),
               L.ptr(self.split_npiece), L.ptr(ws), L.stream())
        self.slot_of_entry = self.slot_of_entry[:nnz]
        torch.cuda.current_stream().synchronize()        # `ws` is released below

    def _build_torch(self, rowptr, col, nnz, num_workers):
        dev = rowptr.device
        n, S = self.num_rows, self.SLOTS
        rp = rowptr.long()
        deg = rp[1:] - rp[:-1]
        nbr = torch.clamp((deg + S - 1) // S, min=1)            # an empty row is one all-padding batch
        bptr = torch.cumsum(nbr, 0) - nbr
        nb = int(nbr.sum().item()) if n else 0
        self.num_rows, self.num_batches = n, nb
        self.num_workers = max(1, min(int(num_workers), max(nb, 1)))
        per = -(-nb // self.num_workers) if nb else 1
        self.batches_per_worker = per
        self.num_workers = -(-nb // per) if nb else 1
        ar = torch.arange(nb, device=dev)
        row_of = torch.repeat_interleave(torch.arange(n, device=dev), nbr)
        k_in = ar - bptr[row_of]
        ebase = rp[row_of] + S * k_in
        slot_e = ebase[:, None] + torch.arange(S, device=dev)[None, :]
        valid = slot_e < rp[row_of + 1][:, None]
        # two batches of slack at the end of desc / colp / every per-slot value buffer: the kernel loads
        # "the next batch" unconditionally
        self.num_slots = (nb + 2) * S
        colp = torch.full((nb + 2, S), -1, dtype=torch.int32, device=dev)
        colp[:nb][valid] = col[:nnz][slot_e[valid]]
        self.colp = colp.reshape(-1).contiguous()
        soe = torch.empty(max(int(nnz), 1), dtype=torch.int64, device=dev)
        soe[slot_e[valid]] = (ar[:, None] * S + torch.arange(S, device=dev)[None, :])[valid]
        self.slot_of_entry = soe[:nnz]
        # ---- flush points: end of a row, or end of a worker's range inside a row
        last_in_row = k_in == nbr[row_of] - 1
        wk = ar // per
        range_end = torch.ones(nb, dtype=torch.bool, device=dev)
        if nb > 1:
            range_end[:-1] = wk[1:] != wk[:-1]
        flush = last_in_row | range_end
        first_b = bptr
        last_b = bptr + nbr - 1
        split = wk[first_b] != wk[last_b] if nb else torch.zeros(0, dtype=torch.bool, device=dev)
        split_rows = split.nonzero().squeeze(1)
        self.num_split = int(split_rows.numel())
        in_split = split[row_of] if nb else split
        piece_flush = flush & in_split
        piece_id = torch.cumsum(piece_flush.long(), 0) - 1
        self.num_piece = int(piece_flush.sum().item()) if nb else 0
        desc = torch.zeros(nb, dtype=torch.int64, device=dev)
        desc = torch.where(flush & ~in_split, row_of + self.FLUSH, desc)
        desc = torch.where(piece_flush, piece_id + self.PIECE + self.FLUSH, desc)
        self.desc = torch.cat([desc, desc.new_zeros(2)]).to(torch.int32).contiguous()
        i32 = dict(dtype=torch.int32, device=dev)
        if self.num_split:
            hid = torch.cumsum(split.long(), 0) - 1                   # row -> split index
            self.piece_split = hid[row_of[piece_flush]].to(torch.int32).contiguous()
            npiece = (wk[last_b[split_rows]] - wk[first_b[split_rows]] + 1)
            self.split_row = split_rows.to(torch.int32).contiguous()
            self.split_npiece = npiece.to(torch.int32).contiguous()
            self.split_piece_beg = (torch.cumsum(npiece, 0) - npiece).to(torch.int32).contiguous()
            self.split_ticket = torch.zeros(self.num_split, **i32)
        else:
            self.piece_split = self.split_row = self.split_npiece = self.split_piece_beg = self.split_ticket = None
    def _finish(self):
        n, nb, per = self.num_rows, self.num_batches, self.batches_per_worker
        self._scratch = {}
        self._cs = {}
        s = L.BplanStruct()
        s.num_rows, s.num_batches = n, nb
        s.num_workers, s.batches_per_worker = self.num_workers, per
        s.desc, s.colp = self.desc.data_ptr(), self.colp.data_ptr()
        s.num_split, s.num_piece = self.num_split, self.num_piece
        if self.num_split:
            for k in ('piece_split', 'split_row', 'split_piece_beg', 'split_npiece', 'split_ticket'):
                setattr(s, k, getattr(self, k).data_ptr())
        self.struct = s
        self.ref = C.byref(s)

    def scratch(self, feat):
        if self.num_piece == 0:
            return None
        buf = self._scratch.get(feat)
        if buf is None:
            buf = torch.empty(self.num_piece * feat, dtype=torch.float32, device=self.desc.device)
            self._scratch[feat] = buf
        return buf

    def col_scale_weights(self, col_scale):
        """Padded per-slot weights ``col_scale[colp]`` of a constant column scale (GCN's D^-1/2 on the
        transpose-backward), cached per scale tensor."""
        key = (col_scale.data_ptr(), col_scale._version)
        hit = self._cs.get(key)
        if hit is None:
            w = torch.zeros(self.num_slots, dtype=torch.float32, device=col_scale.device)
            ok = self.colp >= 0
            w[ok] = col_scale[self.colp[ok].long()]
            self._cs = {key: (w, col_scale)}          # the cached entry keeps the source tensor alive (pointer reuse)
            hit = self._cs[key]
        return hit[0]

    def pad_values(self, val, out=None):
        """Per-entry values (CSR order) -> padded slot layout (padding slots 0)."""
        if out is None:
            out = torch.zeros(self.num_slots, dtype=torch.float32, device=val.device)
        out[self.slot_of_entry] = val[:self.slot_of_entry.numel()]
        return out


def degree_window_perm(rowptr, window=DEG_SORT_WINDOW):
    """Visiting order for the aggregation kernels: rows sorted by degree inside windows of
    ``window`` consecutive rows, so that the 4-8 rows sharing a warp have similar length
    (less intra-warp idling) while coarse row locality is kept."""
    n = rowptr.numel() - 1
    deg = (rowptr[1:] - rowptr[:-1]).long()
    ids = torch.arange(n, device=rowptr.device)
    key = (ids // window) * (int(deg.max().item()) + 1 if n else 1) + deg
    return torch.argsort(key, stable=True).to(torch.int32)


def build_csr(src, dst, num_nodes, self_loops=False, rel=None, num_rel=1, seg_len=SEG_LEN, deg_sort=False):
    """COO (int64, ``src -> dst``) -> :class:`CSR` via ``gd_csr_from_coo`` +
    ``gd_spmm_plan_build``.  Raises on out-of-range endpoints."""
    dev = src.device
    src = src.contiguous()
    dst = dst.contiguous()
    E, N = src.numel(), int(num_nodes)
    cap = E + (N if self_loops else 0)
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    col = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    eid = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    rel_out = torch.empty(max(cap, 1), dtype=torch.int32, device=dev) if rel is not None else None
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    ws_bytes = L.load().gd_csr_workspace_bytes(E, N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    L.call('gd_csr_from_coo', L.ptr(src, 'i64'), L.ptr(dst, 'i64'),
           L.ptr(rel.contiguous(), 'i64') if rel is not None else None,
           E, N, int(num_rel), int(bool(self_loops)), L.ptr(rowptr), L.ptr(col), L.ptr(eid),
           L.ptr(rel_out), L.ptr(status), L.ptr(ws), ws_bytes, L.stream())
    nnz, bad = status.tolist()
    if bad:
        raise ValueError(f'edge_index holds {bad} endpoints / relation ids outside [0, {N})')
    del ws
    col, eid = col[:nnz], eid[:nnz]
    if rel_out is not None:
        rel_out = rel_out[:nnz]
    plan = None
    if seg_len and nnz > 0:
        hcap, scap = nnz // seg_len + 1, 2 * nnz // seg_len + 2
        bufs = {k: torch.empty(hcap if k.startswith('heavy') else scap, dtype=torch.int32, device=dev)
                for k in ('heavy_row', 'heavy_seg_beg', 'heavy_nseg', 'seg_row', 'seg_beg', 'seg_heavy')}
        counts = torch.zeros(2, dtype=torch.int32, device=dev)
        L.call('gd_spmm_plan_build', L.ptr(rowptr), N, int(seg_len), L.ptr(bufs['heavy_row']),
               L.ptr(bufs['heavy_seg_beg']), L.ptr(bufs['heavy_nseg']), L.ptr(bufs['seg_row']),
               L.ptr(bufs['seg_beg']), L.ptr(bufs['seg_heavy']), L.ptr(counts), L.stream())
        nh, ns = counts.tolist()
        if nh > 0:
            bufs['heavy_ticket'] = torch.zeros(nh, dtype=torch.int32, device=dev)
            plan = dict(seg_len=int(seg_len), num_heavy=nh, num_seg=ns, **bufs)
    perm = None      # visiting-order permutation: unused by the row-pipelined kernel (kept in the ABI)
    groups = None    # row groups of the (removed) streaming kernel: the ABI field stays reserved
    return CSR(rowptr, col, eid, rel_out, N, nnz, plan, perm, groups)


def invert_perm(perm, n_out):
    inv = torch.empty(n_out, dtype=torch.int32, device=perm.device)
    L.call('gd_invert_perm', L.ptr(perm, 'i32'), perm.numel(), L.ptr(inv), L.stream())
    return inv


class GraphPlan:
    """Everything a conv layer needs for one fixed edge set.

    ``fwd``: CSR over destinations (messages j -> i).  ``bwd``: CSR of the transposed
    edge set for the gradient w.r.t. the source features (shared with ``fwd`` when the
    edge set is symmetric, which ``to_undirected`` guarantees on the reference's data,
    delete_gnn.py:175-182).  ``dinv``: GCN ``deg^-1/2`` (self loops included)."""

    def __init__(self, edge_index, num_nodes, self_loops, edge_type=None, num_rel=1, gcn_norm=False,
                 need_bwd=True):
        src, dst = edge_index[0], edge_index[1]
        self.num_nodes = int(num_nodes)
        self.self_loops = bool(self_loops)
        self.fwd = build_csr(src, dst, num_nodes, self_loops, edge_type, num_rel)
        self.bwd = None
        self.symmetric = False
        self.bwd_eid = None      # edge ids in TRANSPOSED order (differs from fwd.eid even when symmetric)
        if need_bwd:
            bwd = build_csr(dst, src, num_nodes, self_loops, edge_type, num_rel)
            self.bwd_eid = bwd.eid
            if edge_type is None and bwd.nnz == self.fwd.nnz and torch.equal(bwd.rowptr, self.fwd.rowptr) \
                    and torch.equal(bwd.col, self.fwd.col):
                self.symmetric = True
                self.bwd = self.fwd
            else:
                self.bwd = bwd
        self._tinv = None
        self.dinv = None
        if gcn_norm:
            self.dinv = torch.empty(self.num_nodes, dtype=torch.float32, device=src.device)
            L.call('gd_gcn_dinv', L.ptr(self.fwd.rowptr), self.num_nodes, L.ptr(self.dinv), L.stream())


def _tinv_of(plan):
    """fwd CSR position -> position of the same edge in the transposed CSR (edge-valued
    backward passes write per-edge values straight into transposed order)."""
    if plan._tinv is None:
        n_ids = int(max(plan.fwd.eid.max().item(), plan.bwd_eid.max().item())) + 1 if plan.fwd.nnz else 0
        inv_bwd = invert_perm(plan.bwd_eid, max(n_ids, 1))
        plan._tinv = inv_bwd[plan.fwd.eid.long()].contiguous()
    return plan._tinv


GraphPlan.tinv = property(_tinv_of)


def _rgcn_weights_of(plan):
    """(entry weights of the forward CSR, the same weights in transposed-CSR order):
    1 / |N_r(i)| per entry — RGCNConv's per-(destination, relation) mean."""
    if getattr(plan, '_rgcn_w', None) is None:
        dev = plan.fwd.rowptr.device
        w = torch.empty(max(plan.fwd.nnz, 1), dtype=torch.float32, device=dev)
        L.call('gd_rgcn_norm', L.ptr(plan.fwd.rowptr), L.ptr(plan.fwd.rel), plan.num_nodes, L.ptr(w), L.stream())
        wt = torch.empty_like(w)
        L.call('gd_permute_f32', L.ptr(w), L.ptr(plan.tinv), plan.fwd.nnz, L.ptr(wt), L.stream())
        plan._rgcn_w = (w, wt)
    return plan._rgcn_w


GraphPlan.rgcn_weights = property(_rgcn_weights_of)


def _rgcn_virtual(plan, transposed):
    """CSR over the VIRTUAL sources ``rel * N + col`` (the rows of the stacked per-relation transforms
    ``[R N, F]``) for the transform-then-gather RGCN path; ``None`` when the index would overflow int32."""
    cache = plan.__dict__.setdefault('_rgcn_vcsr', {})
    if transposed not in cache:
        csr = plan.bwd if transposed else plan.fwd
        num_rel = int(csr.rel.max().item()) + 1 if (csr.rel is not None and csr.nnz) else 1
        if csr.rel is None or num_rel * plan.num_nodes >= (1 << 31):
            cache[transposed] = None
        else:
            vcol = (csr.rel.long() * plan.num_nodes + csr.col.long()).to(torch.int32).contiguous()
            cache[transposed] = CSR(csr.rowptr, vcol, csr.eid, None, csr.num_rows, csr.nnz)
    return cache[transposed]


def _rgcn_virtual_weights(plan, transposed, bp):
    """The per-entry mean weights ``1 / |N_r(i)|`` in the padded slot layout of batch plan ``bp``."""
    cache = plan.__dict__.setdefault('_rgcn_vw', {})
    key = (transposed, id(bp))
    if key not in cache:
        w_fwd, w_bwd = plan.rgcn_weights
        cache[key] = bp.pad_values(w_bwd if transposed else w_fwd)
    return cache[key]


GraphPlan.rgcn_virtual = _rgcn_virtual
GraphPlan.rgcn_virtual_weights = _rgcn_virtual_weights


class PlanCache:
    """Two-level cache: tensor identity ``(data_ptr, shape, version)`` first, then a
    content check against the cached copy, so that callers which re-materialise the
    same edge set every epoch (``edge_index[:, sdf_mask]``, gnndelete.py:215) do not
    trigger a rebuild."""

    def __init__(self, max_entries=8):
        self.max_entries = max_entries
        self._by_id = {}
        self._entries = []      # (edge_index_copy, edge_type_copy, key, plan)

    @staticmethod
    def _ident(t):
        return None if t is None else (t.data_ptr(), tuple(t.shape), t._version, str(t.device))

    def get(self, edge_index, edge_type, key, builder):
        ident = (self._ident(edge_index), self._ident(edge_type), key)
        hit = self._by_id.get(ident)
        # an identity hit only counts for the very tensor objects it was recorded for (kept alive by the entry): the
        # caching allocator hands the address of a freed ``ei[:, mask]`` temporary to the next edge set of that shape
        if hit is not None and hit[0] is edge_index and hit[1] is edge_type:
            return hit[2]
        for ei, et, k, plan in self._entries:
            if k == key and ei.shape == edge_index.shape and ei.device == edge_index.device \
                    and torch.equal(ei, edge_index) and (et is None) == (edge_type is None) \
                    and (et is None or (et.shape == edge_type.shape and torch.equal(et, edge_type))):
                self._remember(ident, edge_index, edge_type, plan)
                return plan
        plan = builder()
        self._entries.append((edge_index.clone(), None if edge_type is None else edge_type.clone(), key, plan))
        if len(self._entries) > self.max_entries:
            self._entries.pop(0)
            self._by_id.clear()
        self._remember(ident, edge_index, edge_type, plan)
        return plan

    def _remember(self, ident, edge_index, edge_type, plan):
        if len(self._by_id) >= 16:       # the identity level pins the keyed tensors: keep it small
            self._by_id.clear()
        self._by_id[ident] = (edge_index, edge_type, plan)


_GLOBAL_CACHE = PlanCache()


def plan_for(edge_index, num_nodes, kind, edge_type=None, num_rel=1):
    """``kind``: 'gcn' (self loops + D^-1/2), 'gat' (self loops), 'gin' / 'rgcn' (as is)."""
    self_loops = kind in ('gcn', 'gat')
    key = (kind, int(num_nodes), int(num_rel))
    return _GLOBAL_CACHE.get(
        edge_index, edge_type, key,
        lambda: GraphPlan(edge_index, num_nodes, self_loops, edge_type, num_rel, gcn_norm=(kind == 'gcn')))


def rows_of(mask, num_nodes=None):
    """``(rows, complement)`` int32 index lists of a DeletionLayer mask.  The reference
    indexes with the mask directly (``new_rep[mask]``, deletion.py:25), so a bool
    ``[N]`` mask or an integer index tensor are both accepted."""
    if mask.dtype == torch.bool:
        rows = mask.nonzero().squeeze(1).to(torch.int32)
        comp = (~mask).nonzero().squeeze(1).to(torch.int32)
        return rows, comp
    n = int(num_nodes)
    flag = torch.ones(n, dtype=torch.bool, device=mask.device)
    flag[mask.long()] = False
    return mask.to(torch.int32).contiguous(), flag.nonzero().squeeze(1).to(torch.int32)
