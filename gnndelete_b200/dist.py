"""1-D row partition of the GCNDelete epoch over the GPUs of one node (SURVEY.md §8(e),
BASELINE config 5).  Net-new: the reference has no distributed code (``base.py:21`` single device).

Rank p owns a contiguous block of destination rows ``[lo_p, hi_p)`` — cut so that the per-rank WORK (aggregation
entries + loss incidence entries + a per-row constant) is balanced, not the row count — with their features, CSR
rows, Del masks and the incident-pair lists of the loss.  Every exchanged matrix lives in a SLOT space of
``world * per`` rows (``per`` = largest block, so that ``all_gather_into_tensor`` sees equal blocks):
``slot(g) = rank(g) * per + (g - lo_rank(g))``; CSR columns and pair partners are stored as slots.

Per epoch and rank (``wire`` = bf16 or fp32 blocks on NVLink; the gathered block is also the gather operand):

    forward   H0 = D^-1/2 (X W1^T)                   all-gather [N,128]: ONCE at setup (frozen, input-constant)
              A1_loc = D^-1/2 SpMM(csr_loc, H0) + b1;  Del1;  H1_loc = D^-1/2 (relu(x1) W2^T)
                                                      all-gather -> H1 [N,64]
              A2_loc = D^-1/2 SpMM(csr_loc, H1) + b2;  Del2 -> z_loc
                                                      all-gather -> z  [N,64]
    loss      DEC residuals of this rank's share of the Df items (4 row gathers each) -> coefficients
                                                      all-gather -> 2 n_df floats
              ONE pass over the local nodes' incident pairs (gd_node_loss_fwd_bwd): NI logits recomputed from the
              node's side, dz_loc accumulated; nothing crosses ranks for an NI pair
    backward  dW_del2, dA2_loc                        all-gather -> D^-1/2 dA2 [N,64]
              dH1_loc = D^-1/2 SpMM(csr_loc, .)  (A_hat symmetric), dX1_loc, dW_del1
                                                      all-reduce [dW_del1 | dW_del2 | loss sums] (80 KB)
    Adam      replicated

Per rank the work is 1/world of the single-GPU epoch (every aggregation entry, every incidence entry and every Df
item is processed by exactly one rank); the exchanged volume is 3 [N,64] blocks per epoch.
``PartitionPlan`` is pure index logic (torch, any device) so it is tested on CPU with gloo;
``PartitionedGCNDeleteEngine`` runs the plan with the CUDA kernels.  The three per-epoch halo exchanges are PULLS over
NVLink peer memory (``torch.distributed._symmetric_memory``): every rank publishes its block into a symmetric buffer,
a device-side barrier, then each rank copies the other ranks' blocks with the COPY ENGINES (cudaMemcpyAsync from the
mapped peer buffers) - no SM is spent on the transfer, so a third of the NEXT epoch's layer-1 aggregation (which
depends on nothing this epoch produces) runs underneath each exchange.  (An SM-based NCCL all-gather next to a
memory-bound aggregation just trades time: measured, profiles/r2_scale_config5.md.)  ``exchange='nccl'`` keeps the
all_gather_into_tensor path; the small reductions (DEC coefficients, Del gradients, loss sums) are NCCL either way.
With ``world == 1`` the engine runs without a process group — that is the one-GPU point of the scaling curve, same
kernels, same arithmetic.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops
from .graph import CSR, BatchPlan, build_csr, invert_perm


def nccl_options():
    """Process-group options for the partitioned epoch: NCCL's internal stream at HIGH priority, so that a collective's
    CTAs are dispatched ahead of the aggregation kernels that are enqueued to run next to it."""
    import torch.distributed as dist
    opts = dist.ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = True
    return opts


def row_bounds(num_nodes: int, world: int):
    """Equal row blocks (the unbalanced cut; kept for small graphs and tests)."""
    per = -(-num_nodes // world)
    return [(min(r * per, num_nodes), min((r + 1) * per, num_nodes)) for r in range(world)]


def balanced_bounds(weight: torch.Tensor, world: int):
    """Contiguous row blocks of (nearly) equal total ``weight`` (per-row work estimate, > 0)."""
    n = weight.numel()
    if world == 1:
        return [(0, n)]
    c = torch.cumsum(weight.double(), 0)
    total = float(c[-1])
    targets = torch.tensor([total * r / world for r in range(1, world)], dtype=torch.float64, device=weight.device)
    cuts = torch.searchsorted(c, targets).tolist()
    edges = [0] + [min(max(int(x) + 1, 0), n) for x in cuts] + [n]
    for i in range(1, len(edges)):
        edges[i] = max(edges[i], edges[i - 1])
    return [(edges[r], edges[r + 1]) for r in range(world)]


class PartitionPlan:
    """Everything rank ``rank`` needs, as index tensors.  Names ending in ``_loc`` are local row ids, ``_slot``
    are positions in the gathered (slot) space, everything else is a global node / pair id."""

    def __init__(self, data, neg_edge_index, rank: int, world: int, balance: bool = True, align: int = 8):
        n = int(data.num_nodes)
        self.n, self.rank, self.world = n, rank, world
        dev = data.train_pos_edge_index.device
        ei = data.train_pos_edge_index
        sdf = ei[:, data.sdf_mask]
        df = ei[:, data.df_mask]
        ni = sdf[:, sdf[0] < sdf[1]]                                              # gnndelete.py:379-381
        n_df, n_ni = df.shape[1], ni.shape[1]
        self.norm_df, self.norm_ni = n_df, n_ni
        # ---- row blocks: balance aggregation entries + loss incidence entries + a per-row constant (the row GEMMs)
        keep = sdf[0] != sdf[1]
        if balance and world > 1:
            w = torch.full((n,), 8.0, dtype=torch.float32, device=dev)
            one = torch.ones(1, dtype=torch.float32, device=dev)
            w.index_add_(0, sdf[1][keep], one.expand(int(keep.sum())))
            for t in (ni[0], ni[1], df[0], df[1], neg_edge_index[0], neg_edge_index[1]):
                w.index_add_(0, t, one.expand(t.numel()))
            self.bounds = balanced_bounds(w, world)
        else:
            self.bounds = row_bounds(n, world)
        lo, hi = self.bounds[rank]
        self.lo, self.hi, self.n_loc = lo, hi, hi - lo
        per = max(b[1] - b[0] for b in self.bounds)
        self.per = -(-max(per, 1) // align) * align
        self.num_slots = self.per * world
        self._los = torch.tensor([b[0] for b in self.bounds], dtype=torch.int64, device=dev)
        self._his = torch.tensor([b[1] for b in self.bounds], dtype=torch.int64, device=dev)
        inside = lambda t: (t >= lo) & (t < hi)                                   # noqa: E731
        # ---- message passing: entries with a local destination, PyG self-loop handling (gcn.py:11-12)
        m = inside(sdf[1]) & keep
        loops = torch.arange(lo, hi, device=dev, dtype=ei.dtype)
        self.mp_src_slot = torch.cat([self.slot_of(sdf[0][m]), self.slot_of(loops)])
        self.mp_dst_loc = torch.cat([sdf[1][m] - lo, loops - lo])
        # ---- Del masks (local row ids)
        m1 = data.sdf_node_1hop_mask.to(dev)[lo:hi]
        m2 = data.sdf_node_2hop_mask.to(dev)[lo:hi]
        self.rows1_loc, self.comp1_loc = m1.nonzero().squeeze(1), (~m1).nonzero().squeeze(1)
        self.rows2_loc, self.comp2_loc = m2.nonzero().squeeze(1), (~m2).nonzero().squeeze(1)
        # ---- DEC items (Df pair i with its negative): rank p computes the residuals of items [i_lo, i_hi)
        self.per_items = -(-max(n_df, 1) // world)
        self.i_lo = min(rank * self.per_items, n_df)
        self.i_hi = min((rank + 1) * self.per_items, n_df)
        self.n_items = self.i_hi - self.i_lo
        own = slice(self.i_lo, self.i_hi)
        nu, nv = neg_edge_index[0], neg_edge_index[1]
        self.dec_pu_slot = torch.cat([self.slot_of(df[0][own]), self.slot_of(nu[own])])   # [own Df pairs | their negatives]
        self.dec_pv_slot = torch.cat([self.slot_of(df[1][own]), self.slot_of(nv[own])])
        # ---- incident pairs of the LOCAL nodes.  NI pair j: node <- partner with target j.  DEC pair (i, side):
        #      coefficient = element coef_index(i, side) of the all-gathered [world, 2, per_items] coefficient block
        iu, iv = inside(ni[0]), inside(ni[1])
        jn = torch.arange(n_ni, device=dev)
        self.ni_node_loc = torch.cat([ni[0][iu] - lo, ni[1][iv] - lo])
        self.ni_partner_slot = torch.cat([self.slot_of(ni[1][iu]), self.slot_of(ni[0][iv])])
        self.ni_pair = torch.cat([jn[iu], jn[iv]])
        idf = torch.arange(n_df, device=dev)
        node, partner, cidx = [], [], []
        for side, (a, b) in enumerate(((df[0], df[1]), (nu, nv))):                # side 0: Df pair (+c), side 1: negative (-c)
            for w_, x_ in ((a, b), (b, a)):
                sel = inside(w_)
                node.append(w_[sel] - lo)
                partner.append(self.slot_of(x_[sel]))
                cidx.append(self.coef_index(idf[sel], side))
        self.dec_node_loc = torch.cat(node)
        self.dec_partner_slot = torch.cat(partner)
        self.dec_coef_idx = torch.cat(cidx)

    def slot_of(self, ids):
        """Global node ids -> positions in the gathered slot space."""
        r = torch.searchsorted(self._his, ids, right=True)
        return r * self.per + (ids - self._los[r])

    def coef_index(self, item, side):
        """Position of DEC item ``item``'s coefficient (``side`` 0: Df pair, 1: its negative) in the all-gathered
        ``[world, 2, per_items]`` block."""
        r = torch.div(item, self.per_items, rounding_mode='floor')
        return (r * 2 + side) * self.per_items + (item - r * self.per_items)


class PartitionedGCNDeleteEngine:
    """The epoch of ``engine.GCNDeleteEngine`` on one rank of the row partition."""

    def __init__(self, model, data, neg_edge_index, z_ori_full, group=None, hoist_gather=True, wire='bf16',
                 lr=1e-3, betas=(0.9, 0.999), eps=1e-8, alpha=0.5, balance=True, world=None, rank=None,
                 overlap_layer1=None, exchange='symm'):
        """``wire``: 'bf16' (halo blocks travel and are gathered in bf16, fp32 accumulation; tolerance 2e-2) or
        'fp32' (1e-5).  ``hoist_gather``: the layer-1 block ``H0 = D^-1/2 X W1^T`` is frozen and input-constant, so
        it is transformed and exchanged once at setup; the layer-1 aggregation itself is still run every epoch (the
        reference recomputes conv1 every epoch).  ``world`` / ``rank`` default to the process group's.
        ``exchange``: 'symm' (copy-engine pulls from symmetric peer buffers; falls back to 'nccl' when symmetric memory
        cannot be set up) or 'nccl'.  ``overlap_layer1`` (default: on when ``world > 1``, the H0 exchange is hoisted and
        the exchange is 'symm'): the layer-1 aggregation of epoch k + 1 depends on nothing epoch k produces, so it is
        cut into three row ranges that run underneath the three exchanges of epoch k (double-buffered ``a1``)."""
        if wire not in ('bf16', 'fp32'):
            raise ValueError(wire)
        if world is None:
            import torch.distributed as dist
            world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.world, self.rank, self.group = int(world), int(rank), group
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
        self.model, self.wire = model, wire
        self.wdtype = torch.bfloat16 if wire == 'bf16' else torch.float32
        dev = data.x.device
        plan = PartitionPlan(data, neg_edge_index, self.rank, self.world, balance=balance)
        self.plan = plan
        n, nl, per, slots = plan.n, plan.n_loc, plan.per, plan.num_slots
        self.nl = nl
        if self.world > 1:
            # the ranks must have cut the SAME graph: identical block size and pair counts everywhere
            sig = [plan.per, plan.norm_df, plan.norm_ni, int(data.sdf_mask.sum())]
            chk = torch.tensor(sig + [-v for v in sig], dtype=torch.int64, device=dev)
            self.dist.all_reduce(chk, op=self.dist.ReduceOp.MAX, group=group)
            chk = chk.tolist()
            if any(chk[i] != -chk[i + len(sig)] for i in range(len(sig))):
                raise RuntimeError('the ranks of the partition hold different graphs (block size / pair counts differ): '
                                   'give every rank the same inputs (e.g. broadcast them from rank 0)')
        hid, out = model.conv1.out_channels, model.conv2.out_channels
        self.x_loc = data.x[plan.lo:plan.hi].contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = lambda t: t.to(torch.int32).contiguous()                           # noqa: E731
        # ---- local CSR over source slots (rows beyond n_loc are empty -> truncated view)
        full = build_csr(plan.mp_src_slot, plan.mp_dst_loc, slots, self_loops=False)
        self.csr = _truncate(full, nl)
        self.dinv = torch.empty(max(nl, 1), **f32)
        L.call('gd_gcn_dinv', L.ptr(self.csr.rowptr), nl, L.ptr(self.dinv), L.stream())
        self.rows1, self.comp1 = i32(plan.rows1_loc), i32(plan.comp1_loc)
        self.rows2, self.comp2 = i32(plan.rows2_loc), i32(plan.comp2_loc)
        # ---- buffers: *_send is this rank's [per, F] block in wire format, the gathered matrices are [slots, F]
        self.exchange = exchange if self.world > 1 else 'none'
        self.sym = {}
        self.h0 = self._alloc_gathered('h0', slots, hid, symmetric=False)
        self.h1 = self._alloc_gathered('h1', slots, out)
        self.zg = self._alloc_gathered('z', slots, out)
        self.da2g = self._alloc_gathered('da2', slots, out)
        blk = slice(self.rank * per, self.rank * per + per)
        self.blk = blk
        self.h0_send, self.h1_send, self.z_send, self.da2_send = self.h0[blk], self.h1[blk], self.zg[blk], self.da2g[blk]
        self.h0_loc = torch.empty(nl, hid, **f32) if wire == 'bf16' else self.h0_send[:nl]
        self.h1_loc = torch.empty(nl, out, **f32) if wire == 'bf16' else self.h1_send[:nl]
        self.z_loc = torch.empty(nl, out, **f32) if wire == 'bf16' else self.z_send[:nl]
        self.da2_loc = torch.empty(nl, out, **f32)
        if self.exchange == 'symm':
            self.copy_stream = torch.cuda.Stream()
        want = bool(hoist_gather and self.exchange == 'symm')
        self.overlap = want if overlap_layer1 is None else bool(overlap_layer1 and want)
        self.a1_bufs = [torch.empty(nl, hid, **f32) for _ in range(2 if self.overlap else 1)]
        self.a1 = self.a1_bufs[0]
        self.x1 = torch.empty(nl, hid, **f32)
        self._k, self._a1_pending = 0, False
        if self.overlap:
            # three row ranges of the local CSR, each with its own batch plan: one per exchange of the epoch
            cuts = [0, nl // 3, 2 * nl // 3, nl]
            self.l1_parts = [(cuts[i], cuts[i + 1], _row_range(self.csr, cuts[i], cuts[i + 1])) for i in range(3) if cuts[i + 1] > cuts[i]]
        else:
            self.l1_parts = [(0, nl, self.csr)]
        self.a2 = torch.empty(nl, out, **f32)
        self.dz = torch.empty(nl, out, **f32)
        self.dh1 = torch.empty(nl, out, **f32)
        self.dx1 = torch.zeros(nl, hid, **f32)
        # ---- loss: DEC pre-pass over this rank's items, node pass over the local incidence
        self.alpha = float(alpha)
        ni = plan.n_items
        self.dec_pu, self.dec_pv = i32(plan.dec_pu_slot), i32(plan.dec_pv_slot)
        if ni == 0:
            self.dec_pu = torch.zeros(1, dtype=torch.int32, device=dev); self.dec_pv = self.dec_pu.clone()
        pi = plan.per_items
        self.coef_all = torch.zeros(self.world * 2 * pi, **f32)                  # [world, 2, per_items]
        self.coef_loc = self.coef_all[self.rank * 2 * pi:(self.rank + 1) * 2 * pi]
        self.coef_send = torch.zeros(2 * pi, **f32)
        self.loss_r_part = torch.zeros(1, **f32)
        self.dec_ws_bytes = L.load().gd_dec_items_workspace_bytes()
        self.dec_ws = torch.empty(self.dec_ws_bytes, dtype=torch.uint8, device=dev)
        self._build_incidence(plan, z_ori_full, out)
        # ---- reductions: [dW_del1 | dW_del2 | sum_r, sum_l]
        self.params = [model.deletion1.deletion_weight, model.deletion2.deletion_weight]
        for p in self.params:
            p.grad = torch.zeros_like(p)
        self.n1, self.n2 = self.params[0].numel(), self.params[1].numel()
        self.red = torch.zeros(self.n1 + self.n2 + 2, **f32)
        self.losses = torch.zeros(3, **f32)
        self.node_losses = torch.zeros(3, **f32)
        self.ni_sq = torch.zeros(1, **f32)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.state = [dict(m=torch.zeros_like(p), v=torch.zeros_like(p), step=torch.zeros(1, **f32)) for p in self.params]
        self.hoist_gather, self._h0_done = bool(hoist_gather), False
        self.comm_events = None                                                 # list of (start, end) CUDA events when timing
        self.halo_bytes_per_epoch = 3 * self.zg.numel() * self.zg.element_size() + \
            (0 if hoist_gather else self.h0.numel() * self.h0.element_size())

    def _build_incidence(self, plan, z_ori_full, feat):
        """Batch plan over the local nodes' incident pairs; NI targets in their slots; the (slot, source) index lists
        that drop the all-gathered DEC coefficients into theirs."""
        dev = self.x_loc.device
        nl, slots = plan.n_loc, plan.num_slots
        n_ni_ent = plan.ni_node_loc.numel()
        dst = torch.cat([plan.ni_node_loc, plan.dec_node_loc])
        src = torch.cat([plan.ni_partner_slot, plan.dec_partner_slot])
        bf16 = int(self.wire == 'bf16')
        self.inc_ok = dst.numel() > 0 and nl > 0
        if not self.inc_ok:
            return
        if nl >= (1 << 24):
            raise NotImplementedError('more than 2^24 local rows: split the graph over more ranks')
        inc = _truncate(build_csr(src, dst, slots, self_loops=False), nl)
        pos = invert_perm(inc.eid, max(inc.nnz, 1)).long()[:inc.nnz]
        workers = L.load().gd_node_loss_workers(int(feat), bf16)
        bp = BatchPlan(inc.rowptr, inc.col, nl, inc.nnz, workers)
        bp.colp.clamp_(min=0)       # padding slots gather a valid row (coefficient 0): the kernel's loads are unpredicated
        self.inc, self.inc_bp = inc, bp
        slot = bp.slot_of_entry[pos]
        self.inc_val = torch.zeros(bp.num_slots, dtype=torch.float32, device=dev)
        flag = torch.zeros(bp.num_slots, dtype=torch.bool, device=dev)
        if n_ni_ent:
            # original-model logits of the NI pairs seen from this rank (gnndelete.py:383), once
            gl = plan.lo + plan.ni_node_loc
            part_global = self._global_of_slot(plan, plan.ni_partner_slot)
            tgt = ops.pair_decode(z_ori_full, gl.to(torch.int32).contiguous(), part_global.to(torch.int32).contiguous())
            self.inc_val[slot[:n_ni_ent]] = tgt
            flag[slot[:n_ni_ent]] = True
        self.dec_slots = slot[n_ni_ent:].to(torch.int32).contiguous()
        self.dec_src = plan.dec_coef_idx.to(torch.int32).contiguous()
        deg = (inc.rowptr[1:] - inc.rowptr[:-1]).long()
        nbr = torch.clamp((deg + 7) // 8, min=1)
        brow = torch.repeat_interleave(torch.arange(nl, device=dev), nbr)
        bits = (flag.view(-1, 8).long() << torch.arange(8, device=dev)).sum(1)
        meta = bits << 24
        meta[:bp.num_batches] |= brow
        self.bmeta = torch.where(meta >= (1 << 31), meta - (1 << 32), meta).to(torch.int32).contiguous()
        self.nl_ws_bytes = L.load().gd_node_loss_workspace_bytes(bp.num_workers)
        self.nl_ws = torch.empty(self.nl_ws_bytes, dtype=torch.uint8, device=dev)

    @staticmethod
    def _global_of_slot(plan, slot):
        r = torch.div(slot, plan.per, rounding_mode='floor')
        return plan._los[r] + (slot - r * plan.per)

    # ---------------------------------------------------------------- exchanges
    def _alloc_gathered(self, name, rows, feat, symmetric=True):
        """A [slots, feat] matrix in wire format; with the 'symm' exchange it lives in symmetric memory and every rank
        maps the other ranks' copies (their blocks are pulled from there)."""
        dev = self.x_loc.device
        if self.exchange == 'symm' and symmetric:
            try:
                import torch.distributed._symmetric_memory as sm
                t = sm.empty(rows, feat, dtype=self.wdtype, device=dev)
                hdl = sm.rendezvous(t, self.group if self.group is not None else self.dist.group.WORLD)
                peers = [hdl.get_buffer(q, (rows, feat), self.wdtype) for q in range(self.world)]
                t.zero_()
                self.sym[name] = dict(hdl=hdl, peers=peers, dirty=False)
                return t
            except Exception as exc:               # no peer mapping on this box: every rank fails the same way
                self.exchange = 'nccl'
                self.exchange_error = repr(exc)
                self.sym = {}
        return torch.zeros(rows, feat, dtype=self.wdtype, device=dev)

    def _timed(self, fn, label, stream=None):
        if self.comm_events is None:
            return fn()
        st = stream or torch.cuda.current_stream()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st)
        self.comm_events.append((label, a, b))

    def _before_publish(self, name):
        """The block about to be overwritten may still be being pulled by a slower rank (previous epoch)."""
        e = self.sym.get(name)
        if e is not None and e['dirty']:
            e['hdl'].barrier(channel=1)
            e['dirty'] = False

    def _exchange_begin(self, name, send, full, label):
        """Start the halo exchange of ``full`` (this rank's block ``send`` is already in place).  'symm': device
        barrier, then world - 1 peer pulls on the copy stream (copy engines); the caller may enqueue independent work on
        the current stream before :meth:`_exchange_end`.  'nccl': the all-gather runs here."""
        if self.world == 1:
            return
        e = self.sym.get(name)
        if e is None:
            self._timed(lambda: self.dist.all_gather_into_tensor(full, send, group=self.group), label)
            return
        main = torch.cuda.current_stream()
        e['hdl'].barrier(channel=0)                          # every rank's block is published
        self.copy_stream.wait_stream(main)
        per, peers = self.plan.per, e['peers']

        def pulls():
            with torch.cuda.stream(self.copy_stream):
                for i in range(1, self.world):
                    q = (self.rank + i) % self.world
                    full[q * per:(q + 1) * per].copy_(peers[q][q * per:(q + 1) * per], non_blocking=True)
        self._timed(pulls, label, stream=self.copy_stream)
        e['dirty'] = True

    def _exchange_end(self, name):
        if self.world > 1 and name in self.sym:
            torch.cuda.current_stream().wait_stream(self.copy_stream)

    def _publish(self, loc, send, row_scale=None):
        """Local fp32 rows -> this rank's wire-format block of the gathered matrix."""
        if self.wire == 'bf16':
            ops.cast_bf16(loc, out=send[:self.nl], row_scale=row_scale)
        elif row_scale is not None:
            torch.mul(loc, row_scale.view(-1, 1), out=send[:self.nl])
        elif loc.data_ptr() != send.data_ptr():
            send[:self.nl].copy_(loc)

    # -------------------------------------------------------------------- forward
    def gather_h0(self):
        c1 = self.model.conv1
        ops.gemm_rows(self.x_loc, c1.lin.weight.detach(), True, out=self.h0_loc, out_scale=self.dinv)
        self._publish(self.h0_loc, self.h0_send)
        if self.world > 1:
            self._timed(lambda: self.dist.all_gather_into_tensor(self.h0, self.h0_send, group=self.group), 'allgather_h0')
        self._h0_done = True

    def _layer1_part(self, i, out):
        """Rows [r0, r1) of the layer-1 aggregation A1 = D^-1/2 SpMM(csr_loc, H0) + b1."""
        if i >= len(self.l1_parts):
            return
        r0, r1, csr = self.l1_parts[i]
        ops.spmm(csr, self.h0, out=out[r0:r1], row_scale=self.dinv[r0:r1], bias=self.model.conv1.bias.detach())

    def _layer1_all(self, out):
        for i in range(len(self.l1_parts)):
            self._layer1_part(i, out)

    def forward(self):
        m, nl = self.model, self.nl
        if not (self.hoist_gather and self._h0_done):
            self.gather_h0()
        cur = self._k & 1 if self.overlap else 0
        self.a1 = self.a1_bufs[cur]
        nxt = self.a1_bufs[1 - cur] if self.overlap else None
        if not (self.overlap and self._a1_pending):
            self._layer1_all(self.a1)                              # otherwise: done underneath the previous epoch's exchanges
        w1, w2 = m.deletion1.deletion_weight.detach(), m.deletion2.deletion_weight.detach()
        ops.gemm_rows(self.a1, w1, False, out=self.x1, rows=self.rows1)
        ops.copy_rows(self.a1, self.x1, self.comp1)
        ops.gemm_rows(self.x1, m.conv2.lin.weight.detach(), True, out=self.h1_loc, out_scale=self.dinv, relu_in=True)
        self._before_publish('h1')
        self._publish(self.h1_loc, self.h1_send)
        self._exchange_begin('h1', self.h1_send, self.h1, 'exchange_h1')
        if self.overlap:
            self._layer1_part(0, nxt)
        self._exchange_end('h1')
        ops.spmm(self.csr, self.h1, out=self.a2, row_scale=self.dinv, bias=m.conv2.bias.detach())
        ops.gemm_rows(self.a2, w2, False, out=self.z_loc, rows=self.rows2)
        ops.copy_rows(self.a2, self.z_loc, self.comp2)
        self._before_publish('z')
        self._publish(self.z_loc, self.z_send)
        self._exchange_begin('z', self.z_send, self.zg, 'exchange_z')
        if self.overlap:
            self._layer1_part(1, nxt)
        self._exchange_end('z')
        self._loss()

    def _loss(self):
        p, bf16 = self.plan, int(self.wire == 'bf16')
        z, pi = self.zg, self.plan.per_items
        send = self.coef_send if self.world > 1 else self.coef_loc
        L.call('gd_dec_items_fwd', z.data_ptr(), z.stride(0), bf16, z.shape[1], L.ptr(self.dec_pu), L.ptr(self.dec_pv),
               p.n_items, p.norm_df, self.alpha, L.ptr(send), L.ptr(send[pi:]), None, L.ptr(self.loss_r_part),
               L.ptr(self.dec_ws), self.dec_ws_bytes, L.stream())
        if self.world > 1:
            self._timed(lambda: self.dist.all_gather_into_tensor(self.coef_all, self.coef_send, group=self.group), 'allgather_dec_coef')
        if not self.inc_ok:
            self.dz.zero_()
            self.ni_sq.zero_()
            return
        if self.dec_slots.numel():
            L.call('gd_move_f32', L.ptr(self.coef_all), L.ptr(self.dec_src), L.ptr(self.inc_val), L.ptr(self.dec_slots),
                   self.dec_slots.numel(), L.stream())
        bp = self.inc_bp
        zself = z[self.rank * p.per:]
        L.call('gd_node_loss_fwd_bwd', bp.ref, L.ptr(self.bmeta), L.ptr(self.inc_val), None, None, None,
               z.data_ptr(), z.stride(0), zself.data_ptr(), z.stride(0), bf16, None, z.shape[1], p.norm_ni, self.alpha,
               None, L.ptr(self.dz), self.dz.stride(0), L.ptr(bp.scratch(z.shape[1])), L.ptr(self.node_losses),
               L.ptr(self.ni_sq), L.ptr(self.nl_ws), self.nl_ws_bytes, L.stream())

    # ------------------------------------------------------------------- backward
    def backward(self):
        m = self.model
        g1, g2 = self.params[0].grad, self.params[1].grad
        w2 = m.deletion2.deletion_weight.detach()
        ops.gemm_rows(self.dz, w2, True, out=self.da2_loc, rows=self.rows2)
        ops.copy_rows(self.dz, self.da2_loc, self.comp2)
        self._before_publish('da2')
        self._publish(self.da2_loc, self.da2_send, row_scale=self.dinv)          # D^-1/2 on the source side of A_hat^T
        self._exchange_begin('da2', self.da2_send, self.da2g, 'exchange_da2')
        ops.gemm_tn_rows(self.a2, self.dz, rows=self.rows2, out=g2)             # dW_del2 needs no halo: under the exchange
        if self.overlap:
            self._layer1_part(2, self.a1_bufs[1 - (self._k & 1)])
            self._a1_pending = True
        self._exchange_end('da2')
        ops.spmm(self.csr, self.da2g, out=self.dh1)                              # A^T = A (symmetric edge set)
        ops.gemm_rows(self.dh1, m.conv2.lin.weight.detach(), False, out=self.dx1, rows=self.rows1,
                      out_scale=self.dinv, gate=self.x1)
        ops.gemm_tn_rows(self.a1, self.dx1, rows=self.rows1, out=g1)
        n1, n2 = self.n1, self.n2
        self.red[:n1].copy_(g1.view(-1)); self.red[n1:n1 + n2].copy_(g2.view(-1))
        # loss sums: loss_r_part = (sum of this rank's squared DEC residuals) / norm_df; ni_sq = raw NI sum (x2)
        self.red[n1 + n2:n1 + n2 + 1].copy_(self.loss_r_part)
        self.red[n1 + n2 + 1:].copy_(self.ni_sq)
        if self.world > 1:
            self._timed(lambda: self.dist.all_reduce(self.red, group=self.group), 'allreduce_grads_losses')
        g1.view(-1).copy_(self.red[:n1]); g2.view(-1).copy_(self.red[n1:n1 + n2])
        loss_r = self.red[n1 + n2]
        loss_l = self.red[n1 + n2 + 1] * (0.5 / self.plan.norm_ni if self.plan.norm_ni else 0.0)
        torch.stack([self.alpha * loss_r + (1.0 - self.alpha) * loss_l, loss_r, loss_l], out=self.losses)
        self._k += 1

    def adam_step(self):
        for p, st in zip(self.params, self.state):
            ops.adam_step(p.data, p.grad, st['m'], st['v'], st['step'], self.lr, self.betas[0], self.betas[1], self.eps)

    def epoch(self):
        self.forward()
        self.backward()
        self.adam_step()
        return self.losses

    def comm_ms(self):
        """{collective: total device ms} since ``comm_events`` was set to a list (synchronises).  The time of a
        collective on one rank includes waiting for the slowest rank to arrive."""
        torch.cuda.synchronize()
        out = {}
        for label, a, b in (self.comm_events or []):
            out[label] = out.get(label, 0.0) + a.elapsed_time(b)
        return out


def _row_range(csr: CSR, r0: int, r1: int) -> CSR:
    """Rows [r0, r1) of a CSR as a CSR of its own (shares the column array)."""
    rp = csr.rowptr[r0:r1 + 1]
    base = int(rp[0].item())
    end = int(rp[-1].item())
    return CSR((rp - base).contiguous(), csr.col[base:end], csr.eid[base:end] if csr.eid is not None else None, None,
               r1 - r0, end - base)


def _truncate(csr: CSR, num_rows: int) -> CSR:
    """View of a CSR whose rows beyond ``num_rows`` are empty."""
    out = CSR(csr.rowptr[:num_rows + 1].contiguous(), csr.col, csr.eid, csr.rel, num_rows, csr.nnz,
              csr.plan if csr.plan else None, None, None)
    return out
