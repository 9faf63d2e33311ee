"""Multi-GPU propagation: the hierarchy is split into contiguous ranges of the
lexicographic storage order, one range per rank (one process per GPU), and the
neighbour rows that cross a range boundary are exchanged once per RK stage.

What the reference does: nothing - it is a single Python loop
(``pyqed/heom/deom.py:1072-1114``).  SURVEY.md section 8e describes the shard:
owner-computes over a partition of the ADOs plus one halo exchange per stage.

Layout per rank: all four ADO arrays keep the full ``[nmax, N, N]`` index space
(links hold global slot numbers, the stage kernel is unchanged) but a rank only
advances its own slots ``[lo, hi)``; the halo consists of *items*: for diagonal
coupling operators one matrix row of a foreign ADO (``16 N`` bytes, the only
part the element-wise coupling needs), otherwise the whole ADO.

The exchange logic (``HaloPlan``) is independent of CUDA so that it is tested
with ``gloo`` on CPU; the device work (stage kernel, pack/unpack kernels) goes
through the C ABI.
"""
from __future__ import annotations

import numpy as np
import torch

C128 = np.complex128
STAGE_OUTPUT_ARRAY = (1, 2, 1, 0)  # array written by RK4 stage 0..3 (include/pyqed_heom.h)


def cost_balanced_bounds(link_ptr, world, base_cost=12):
    """Range boundaries ``b[0]=0 <= ... <= b[world]=nmax`` that equalise
    ``sum(base_cost + links)`` per rank; ``link_ptr`` is the CSR offset array."""
    lp = np.asarray(link_ptr, dtype=np.int64)
    nmax = len(lp) - 1
    cost = base_cost * np.arange(nmax + 1, dtype=np.int64) + (lp - lp[0])
    targets = cost[-1] * np.arange(1, world) / world
    inner = np.searchsorted(cost, targets).astype(np.int64)
    return [0] + [int(x) for x in inner] + [nmax]


def needed_items(nbr, meta, lo, hi, row_items, supp_rows=None):
    """Sorted unique halo items of the links ``(nbr, meta)`` of the owned slots
    ``[lo, hi)``.  ``row_items``: ``slot*8 + row`` for every row in the support
    of the link's coupling mode (``supp_rows[mode]``; default: the single row
    stored in bits 16-19 of ``meta``); otherwise just ``slot``."""
    nbr = nbr.to(torch.int64)
    outside = (nbr < lo) | (nbr >= hi)
    nb = nbr[outside]
    if not row_items:
        return torch.unique(nb)
    mt = meta[outside].to(torch.int64)
    if supp_rows is None:
        return torch.unique(nb * 8 + ((mt >> 16) & 0xF))
    mode = (mt >> 24) & 0xFF
    parts = []
    for m, rows in enumerate(supp_rows):
        sel = nb[mode == m]
        for r in rows:
            parts.append(sel * 8 + int(r))
    if not parts:
        return torch.zeros(0, dtype=torch.int64, device=nbr.device)
    return torch.unique(torch.cat(parts))


class HaloPlan:
    """Who sends which items to whom.  ``need`` is this rank's sorted item list;
    construction is collective over ``transport`` (see ``DistTransport`` /
    ``LocalTransport``)."""

    def __init__(self, bounds, rank, world, need, row_items, transport):
        self.bounds, self.rank, self.world, self.row_items = list(bounds), rank, world, row_items
        self.need = need.to(torch.int64)
        slots = (self.need >> 3) if row_items else self.need
        upper = torch.tensor(self.bounds[1:], dtype=torch.int64, device=slots.device)
        owner = torch.searchsorted(upper, slots, right=True)
        self.recv_counts = [int(x) for x in torch.bincount(owner, minlength=world).tolist()]
        assert self.recv_counts[rank] == 0, "a rank never needs its own slots"
        # counts[q][r] = number of items rank q needs from rank r
        all_counts = transport.allgather_counts(self.recv_counts)
        self.all_counts = [[int(x) for x in row] for row in all_counts]
        self.send_counts = [int(all_counts[q][rank]) for q in range(world)]
        if hasattr(transport, "allgather_lists"):
            # one all-gather of the (sorted) need lists instead of pairwise sends: every rank cuts
            # the part that falls into its own range out of each peer's list.  (Pairwise NCCL
            # sends set up a connection per pair of ranks - seconds at 8 ranks.)
            lists = transport.allgather_lists(self.need, [sum(row) for row in self.all_counts])
            lo_item = self.bounds[rank] * (8 if row_items else 1)
            hi_item = self.bounds[rank + 1] * (8 if row_items else 1)
            got = []
            for q in range(world):
                cut = torch.searchsorted(lists[q], torch.tensor([lo_item, hi_item], dtype=torch.int64,
                                                                device=lists[q].device))
                got.append(lists[q][int(cut[0]):int(cut[1])].to(self.need.device))
                assert got[-1].numel() == self.send_counts[q]
        else:
            chunks = torch.split(self.need, self.recv_counts)
            got = transport.exchange_lists(list(chunks), self.send_counts)
        self.send_items = (torch.cat(got) if sum(self.send_counts) else
                           torch.zeros(0, dtype=torch.int64, device=self.need.device))
        lo, hi = self.bounds[rank], self.bounds[rank + 1]
        s = (self.send_items >> 3) if row_items else self.send_items
        assert bool(((s >= lo) & (s < hi)).all()), "asked for items this rank does not own"


class DistTransport:
    """torch.distributed transport.  NCCL moves device buffers directly
    (``all_to_all_single`` over NVLink); with gloo (CPU tests, or several ranks
    sharing one GPU) buffers are staged through host memory and sent pairwise."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.nccl = self.backend == "nccl"

    def allgather_counts(self, counts):
        mine = torch.tensor(counts, dtype=torch.int64, device="cuda" if self.nccl else "cpu")
        out = [torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(out, mine, group=self.group)
        return [o.tolist() for o in out]

    def exchange_lists(self, chunks, send_counts):
        dev = chunks[0].device
        wire = dev if self.nccl else torch.device("cpu")
        send = [c.to(wire).contiguous() for c in chunks]
        recv = [torch.zeros(n, dtype=torch.int64, device=wire) for n in send_counts]
        self._p2p(send, recv)
        return [r.to(dev) for r in recv]

    def _p2p(self, send, recv):
        ops = []
        for q in range(self.world):
            if q == self.rank:
                continue
            if recv[q].numel():
                ops.append(self.dist.P2POp(self.dist.irecv, recv[q], q, self.group))
            if send[q].numel():
                ops.append(self.dist.P2POp(self.dist.isend, send[q], q, self.group))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def all_to_all(self, sendbuf, send_counts, recvbuf, recv_counts, elems):
        """Buffers are flat float64; counts are in items of ``elems`` reals."""
        ss = [c * elems for c in send_counts]
        rs = [c * elems for c in recv_counts]
        if self.nccl:
            self.dist.all_to_all_single(recvbuf, sendbuf, rs, ss, group=self.group)
            return
        host_send = sendbuf.cpu()
        host_recv = torch.empty(recvbuf.shape, dtype=recvbuf.dtype)
        self._p2p(list(torch.split(host_send, ss)), list(torch.split(host_recv, rs)))
        recvbuf.copy_(host_recv)

    def allgather_lists(self, mine, sizes):
        """Every rank's int64 list (``sizes[q]`` entries from rank q) on every rank."""
        wire = mine.device if self.nccl else torch.device("cpu")
        cap = max(max(sizes), 1)
        buf = torch.zeros(cap, dtype=torch.int64, device=wire)
        buf[:mine.numel()] = mine.to(wire)
        out = [torch.empty(cap, dtype=torch.int64, device=wire) for _ in range(self.world)]
        self.dist.all_gather(out, buf, group=self.group)
        return [o[:n] for o, n in zip(out, sizes)]

    def allgather_object(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def barrier(self):
        self.dist.barrier(group=self.group)

    def broadcast(self, t, src):
        if self.nccl or not t.is_cuda:
            self.dist.broadcast(t, src, group=self.group)
            return
        h = t.cpu()
        self.dist.broadcast(h, src, group=self.group)
        t.copy_(h)

    def allreduce_sum(self, t):
        self.dist.all_reduce(t, group=self.group)


class ShardedDEOM:
    """One rank's share of a sharded RK4 propagation.

    Parameters are those of ``DEOMSolver`` given as arrays; ``transport`` is a
    ``DistTransport`` (one process per GPU).  Only ``batch = 1``.
    """

    def __init__(self, system, system_dipole, coupling, coupling_dipole, expn, etal, etar, etaa,
                 mode, lmax, transport, device=0, order=1, options=None, tuning=None, peer_push=None,
                 fused_push=None, native=None, rebalance=None):
        from .._cabi import Plan
        self.tr = transport
        self.rank, self.world = transport.rank, transport.world
        n = np.shape(system)[0]
        m = int(np.max(mode)) + 1
        self.n = n
        self.device = device
        self.plan = p = Plan(n, len(expn), m, lmax, batch=1, device=device, order=order)
        p.set_system(system, system_dipole)
        p.set_coupling(np.asarray(coupling)[:m], coupling_dipole)
        p.set_bath(expn, etal, etar, etaa, mode)
        if tuning:
            p.set_tuning(**tuning)
        for k, v in (options or {}).items():
            p.set_option(k, v)
        # Rank-local arrays + fused peer stores (csrc/heom_shard.cu) for the problems kernels 6 / 7
        # take; ``native``: None = where it applies, True = required, False = never.
        self.native = False
        self.symm = None
        self.fused = False
        if native is not False and hasattr(transport, "allgather_object"):
            p.build(tables_only=True)
            if p.info("off_links2") >= 0 and p.info("sym_inputs") == 1 and order == 2:
                self._init_native(p)
                # the ranges are first balanced by a static cost (links per ADO); the time of a stage
                # also depends on how many rows a rank stores to and reads from its peers, so large
                # hierarchies are measured once and re-cut (None = from 2^20 ADOs per job)
                if (rebalance if rebalance is not None else (self.nmax >= (1 << 20))) and self.world > 1:
                    self._rebalance(p, force=(rebalance == "force"))
                return
            if native:
                raise ValueError("the rank-local sharded layout needs storage order 2 and a problem for "
                                 "kernels 6 / 7 (Hermitian operators and bath, one-entry diagonal Q_m)")
            p.close()
            self.plan = p = Plan(n, len(expn), m, lmax, batch=1, device=device, order=order)
            p.set_system(system, system_dipole)
            p.set_coupling(np.asarray(coupling)[:m], coupling_dipole)
            p.set_bath(expn, etal, etar, etaa, mode)
            if tuning:
                p.set_tuning(**tuning)
            for k, v in (options or {}).items():
                p.set_option(k, v)
        elif native:
            raise ValueError("the rank-local sharded layout needs a multi-process transport")
        # peer_push: None = use it when NCCL + symmetric memory are available
        self.symm = None
        state = None
        if peer_push is not False and getattr(transport, "nccl", False):
            try:
                import torch.distributed._symmetric_memory as symm_mem
                state = symm_mem.empty(p.state_bytes(), dtype=torch.uint8,
                                       device=torch.device("cuda", device))
                self.symm = symm_mem.rendezvous(state, transport.group or transport.dist.group.WORLD)
            except Exception as exc:  # noqa: BLE001 - fall back to the NCCL exchange
                if peer_push:
                    raise
                self.symm, state, self.symm_error = None, None, repr(exc)
        p.build(state)
        self.nmax = p.nmax
        # a diagonal Q_m only reads the rows r of a neighbour with (Q_m)_rr != 0, and
        # (for Hermitian ADOs) gets the column entries as their conjugates
        # ... which only the row kernels (1, 3) do; the generic kernel reads whole neighbours.  Row items
        # are slot*8+row in 32 bits.
        self.row_items = (bool(p.info("qdiag")) and bool(p.info("hermitian_inputs"))
                          and p.info("stage_kernel") in (1, 3) and self.nmax * 8 < 2 ** 31)
        Qa = np.asarray(coupling, dtype=C128)[:m]
        Qd = (np.zeros_like(Qa) if coupling_dipole is None
              else np.broadcast_to(np.asarray(coupling_dipole, dtype=C128), Qa.shape))
        self.supp_rows = [[r for r in range(n) if Qa[mm, r, r] != 0 or Qd[mm, r, r] != 0]
                          for mm in range(m)]
        tables = p._tables
        lp_off, lk_off = p.info("off_link_ptr"), p.info("off_links")
        self.link_ptr = tables[lp_off:lp_off + 4 * (self.nmax + 1)].view(torch.int32)
        nlinks = p.info("nlinks")
        links = tables[lk_off:lk_off + 8 * nlinks].view(torch.int32).view(nlinks, 2)
        self.bounds = cost_balanced_bounds(self.link_ptr.cpu().numpy(), self.world)
        if order == 2:  # keep the 64-slot blocks of storage order 2 whole
            self.bounds = ([0] + [min(self.nmax, (b + 32) // 64 * 64) for b in self.bounds[1:-1]]
                           + [self.nmax])
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        p._check(p.lib.pyqed_heom_set_partition(p._h, self.lo, self.hi))
        l0, l1 = int(self.link_ptr[self.lo]), int(self.link_ptr[self.hi])
        need = needed_items(links[l0:l1, 0], links[l0:l1, 1], self.lo, self.hi, self.row_items,
                            self.supp_rows)
        self.halo = HaloPlan(self.bounds, self.rank, self.world, need, self.row_items, transport)
        dev = tables.device
        self.elems = 2 * (n if self.row_items else n * n)  # reals per item
        self.need32 = self.halo.need.to(torch.int32).contiguous()
        self.send32 = self.halo.send_items.to(torch.int32).contiguous()
        self.sendbuf = torch.empty(self.send32.numel() * self.elems, dtype=torch.float64, device=dev)
        self.recvbuf = torch.empty(self.need32.numel() * self.elems, dtype=torch.float64, device=dev)
        if self.symm is not None:
            import ctypes as C
            offs = np.concatenate([[0], np.cumsum(self.halo.send_counts)]).astype(np.int64)
            self._push_offs = offs
            self._push_ptrs = (C.c_uint64 * self.world)(*[int(x) for x in self.symm.buffer_ptrs])
        # fused push: the stage kernel itself stores the rows into the peers' arrays
        self.fused = False
        kern = (tuning or {}).get("kernel", 0)
        # (opt-in: measured slower than the separate push kernel in round 1)
        if self.symm is not None and fused_push is True and bool(p.info("rk_scheme")):
            import ctypes as C
            si = self.halo.send_items
            dest = torch.repeat_interleave(
                torch.arange(self.world, device=si.device),
                torch.tensor(self.halo.send_counts, device=si.device))
            if self.row_items:
                loc, rowc = (si >> 3) - self.lo, si & 7
            else:
                loc, rowc = si - self.lo, torch.full_like(si, 15)
            ent = (dest * 16 + rowc).to(torch.uint8)
            order_ix = torch.argsort(loc, stable=True)
            counts = torch.bincount(loc, minlength=self.hi - self.lo)
            ptr = torch.zeros(self.hi - self.lo + 1, dtype=torch.int32, device=si.device)
            ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
            self._push_ptr = ptr.contiguous()
            self._push_ent = ent[order_ix].contiguous() if ent.numel() else torch.zeros(1, dtype=torch.uint8, device=si.device)
            p._check(p.lib.pyqed_heom_set_push_table(p._h, C.c_void_p(self._push_ptr.data_ptr()),
                                                     C.c_void_p(self._push_ent.data_ptr()),
                                                     self._push_ptrs, self.world))
            self.fused = True
        self.owner_of_sys = next(r for r in range(self.world)
                                 if self.bounds[r] <= p.info("slot0") < self.bounds[r + 1])

    # -- rank-local layout: own ADOs + pool of halo rows, rows stored by their owners --------
    def _rebalance(self, p, steps=2, force=False):
        """Re-cut the ranges from measured stage-kernel times: a rank that took longer per ADO gets
        fewer ADOs (cost density assumed uniform inside a rank's old range)."""
        n = self.n
        rho = np.eye(n, dtype=C128) / n
        self.set_state(rho)
        self._native_propagate(1e-9, 1, None)            # warm-up
        p.synchronize()
        p.stage_timing(True)
        self._native_propagate(1e-9, steps, None)
        ms, cnt = p.stage_timing(False)
        mine = ms / max(cnt, 1)
        times = [x[0] / 1e6 for x in self.tr.allgather_counts([int(mine * 1e6)])]   # ms per stage, per rank
        owned = [self.bounds[r + 1] - self.bounds[r] for r in range(self.world)]
        if min(owned) <= 0 or min(times) <= 0 or (max(times) / min(times) < 1.03 and not force):
            self.timings["rebalance"] = {"stage_ms_before": times, "changed": False}
            return
        # cumulative cost over slots, piecewise linear; new boundaries at equal shares
        edges = np.array(self.bounds, dtype=np.float64)
        cum = np.concatenate([[0.0], np.cumsum(times)])
        targets = cum[-1] * np.arange(1, self.world) / self.world
        inner = np.interp(targets, cum, edges)
        new = [0] + [int(min(self.nmax, max(0, round(b / 64.0) * 64))) for b in inner] + [self.nmax]
        if any(b1 <= b0 for b0, b1 in zip(new, new[1:])):
            self.timings["rebalance"] = {"stage_ms_before": times, "changed": False}
            return
        before = dict(self.timings)
        self._release_native()
        self._init_native(p, bounds=new)   # (the rank-local link table is rewritten from the general one)
        self.timings["first_cut"] = before
        self.timings["rebalance"] = {"stage_ms_before": times, "changed": True, "old_bounds": [int(b) for b in edges],
                                     "new_bounds": new}

    def _release_native(self):
        p = self.plan
        self.tr.barrier()
        for q, ptr in enumerate(self._peer_ptrs):
            if q != self.rank:
                p.lib.pyqed_heom_shared_close(self.device, ptr)
        p.lib.pyqed_heom_shared_free(self.device, self._state_ptr)
        self._state_ptr = None
        self.native = False

    def _init_native(self, p, bounds=None):
        import ctypes as C
        import time
        tr, world, rank = self.tr, self.world, self.rank
        self.timings = tm = {}
        t_last = [time.perf_counter()]

        def lap(name):
            torch.cuda.synchronize()
            now = time.perf_counter()
            tm[name] = now - t_last[0]
            t_last[0] = now
        self.nmax = p.nmax
        tables = p._tables
        dev = tables.device
        lp_off, lk_off = p.info("off_link_ptr"), p.info("off_links")
        self.link_ptr = tables[lp_off:lp_off + 4 * (self.nmax + 1)].view(torch.int32)
        nlinks = p.info("nlinks")
        links = tables[lk_off:lk_off + 8 * nlinks].view(torch.int32).view(nlinks, 2)
        if bounds is None:
            self.bounds = cost_balanced_bounds(self.link_ptr.cpu().numpy(), world)
            self.bounds = ([0] + [min(self.nmax, (b + 32) // 64 * 64) for b in self.bounds[1:-1]] + [self.nmax])
        else:
            self.bounds = list(bounds)
        self.lo, self.hi = lo, hi = self.bounds[rank], self.bounds[rank + 1]
        n_own = hi - lo
        lap("bounds")
        l0, l1 = int(self.link_ptr[lo]), int(self.link_ptr[hi])
        need = needed_items(links[l0:l1, 0], links[l0:l1, 1], lo, hi, True)
        self.row_items = True
        lap("need_list")
        self.halo = h = HaloPlan(self.bounds, rank, world, need, True, tr)
        lap("request_exchange")
        self.need64 = h.need.contiguous()
        # push table: for every row a peer q asked for, its index in q's pool = position in q's
        # sorted need list = (rows q gets from lower ranks) + position inside q's request to me
        si = h.send_items
        counts = torch.tensor(h.send_counts, dtype=torch.int64, device=si.device)
        dest = torch.repeat_interleave(torch.arange(world, device=si.device), counts)
        first = torch.cumsum(counts, 0) - counts
        base = torch.tensor([sum(h.all_counts[q][:rank]) for q in range(world)], dtype=torch.int64, device=si.device)
        dst_row = base[dest] + torch.arange(si.numel(), device=si.device) - first[dest]
        loc, rowc = (si >> 3) - lo, si & 7
        # entries ordered by (owned slot, row, destination); every distinct (slot, row) inside a group
        # of 32/N consecutive slots (= the ADOs one warp processes together) gets a staging slot
        key = loc * 8 + rowc
        order_ix = torch.argsort(key * 16 + dest, stable=True)
        key_s, dest_s, dst_s = key[order_ix], dest[order_ix], dst_row[order_ix]
        cnt = torch.bincount(loc, minlength=n_own) if si.numel() else torch.zeros(n_own, dtype=torch.int64, device=si.device)
        ptr = torch.zeros(n_own + 1, dtype=torch.int32, device=si.device)
        ptr[1:] = torch.cumsum(cnt, 0).to(torch.int32)
        if si.numel():
            apw, nslots = 32 // self.n, p.info("push_slots")
            new = torch.ones_like(key_s)
            new[1:] = (key_s[1:] != key_s[:-1]).to(new.dtype)
            c = torch.cumsum(new, 0) - 1
            grp = (key_s >> 3) // apw
            gstart = torch.ones_like(grp, dtype=torch.bool)
            gstart[1:] = grp[1:] != grp[:-1]
            base = c[gstart][torch.cumsum(gstart.to(torch.int64), 0) - 1]
            slot_id = torch.clamp(c - base, max=255)
            slot_id = torch.where(slot_id >= nslots, torch.full_like(slot_id, 255), slot_id)
            ent = torch.stack([dst_s, slot_id * 256 + dest_s * 16 + (key_s & 7)], dim=1).to(torch.int32)
        else:
            ent = torch.zeros((0, 2), dtype=torch.int32, device=si.device)
        self._push_ptr = ptr.to(dev).contiguous()
        self._push_ent = (ent.to(dev).contiguous() if ent.numel()
                          else torch.zeros((1, 2), dtype=torch.int32, device=dev))
        lap("push_table")
        sizes = tr.allgather_counts([n_own, int(self.need64.numel()), int(self.device)])
        n_own_max, pool_max = max(x[0] for x in sizes), max(x[1] for x in sizes)
        self.device_barrier = len({x[2] for x in sizes}) == world   # every rank on its own GPU
        sb, fo = C.c_size_t(), C.c_size_t()
        p._check(p.lib.pyqed_heom_shard_state_bytes(p._h, n_own_max, pool_max, C.byref(sb), C.byref(fo)))
        self.state_nbytes = sb.value
        ptr_own, handle = C.c_void_p(), (C.c_uint8 * 64)()
        p._check(p.lib.pyqed_heom_shared_alloc(self.device, sb.value, C.byref(ptr_own), handle))
        self._state_ptr = ptr_own.value
        handles = tr.allgather_object(bytes(handle))
        self._peer_ptrs = []
        for q in range(world):
            if q == rank:
                self._peer_ptrs.append(self._state_ptr)
                continue
            hq, pq = (C.c_uint8 * 64).from_buffer_copy(handles[q]), C.c_void_p()
            p._check(p.lib.pyqed_heom_shared_open(self.device, hq, C.byref(pq)))
            self._peer_ptrs.append(pq.value)
        peers = (C.c_uint64 * world)(*self._peer_ptrs)
        lap("state_alloc_and_ipc")
        need_dev = self.need64.to(dev)
        self._keep = (need_dev,)
        p._check(p.lib.pyqed_heom_shard_setup(
            p._h, rank, world, lo, hi, n_own_max, pool_max, C.c_void_p(need_dev.data_ptr()), need_dev.numel(),
            C.c_void_p(self._push_ptr.data_ptr()), C.c_void_p(self._push_ent.data_ptr()), int(si.numel()),
            C.c_void_p(self._state_ptr), sb.value, peers, int(self.device_barrier)))
        self.native = True
        self.fused = True
        self.elems = 2 * self.n
        self.owner_of_sys = next(r for r in range(world) if self.bounds[r] <= p.info("slot0") < self.bounds[r + 1])
        tr.barrier()   # every rank's flags are zeroed and its buffer mapped before anyone pushes
        lap("localize_links_and_barrier")

    def close(self):
        """Unmap the peers' buffers and free this rank's (rank-local layout only)."""
        if getattr(self, "plan", None) is None:
            return
        if self.native and getattr(self, "_state_ptr", None):
            self._release_native()
        self.plan.close()
        self.plan = None

    def _host_barrier(self):
        self.plan.synchronize()
        self.tr.barrier()

    def _native_propagate(self, dt, nt, traj):
        import ctypes as C
        p = self.plan
        tp = None if traj is None else C.c_void_p(traj.data_ptr())
        if self.device_barrier:
            p._check(p.lib.pyqed_heom_shard_propagate(p._h, float(dt), int(nt), tp))
            return
        # ranks sharing a GPU: a spinning barrier kernel would keep the peer's kernels off the
        # device, so the ranks meet on the host after every stage
        p._check(p.lib.pyqed_heom_shard_begin(p._h, float(dt), int(nt), tp))
        self._host_barrier()
        for i in range(nt):
            for st in range(4):
                p._check(p.lib.pyqed_heom_shard_stage(p._h, i, st))
                self._host_barrier()
        p._check(p.lib.pyqed_heom_shard_end(p._h))

    def check_barriers(self):
        import ctypes as C
        code = C.c_int()
        self.plan._check(self.plan.lib.pyqed_heom_shard_error(self.plan._h, C.byref(code)))
        if code.value:
            raise RuntimeError(f"rank {self.rank}: a device barrier gave up waiting for rank {code.value - 1}")

    # -- one halo exchange of array `array_id` -----------------------------
    def exchange(self, array_id):
        import ctypes as C
        p = self.plan
        if self.fused:
            # the stage kernel already stored the rows into the peers' arrays
            self.symm.barrier(channel=0)
            return
        if self.symm is not None:
            # store the requested rows straight into the peers' arrays, then a
            # device-side barrier over the symmetric-memory signal pads
            p._check(p.lib.pyqed_heom_halo_push(
                p._h, array_id, C.c_void_p(self.send32.data_ptr()), self.send32.numel(),
                int(self.row_items), self._push_offs.ctypes.data_as(C.POINTER(C.c_int64)),
                self._push_ptrs, self.world))
            self.symm.barrier(channel=0)
            return
        p._check(p.lib.pyqed_heom_halo_pack(p._h, array_id, C.c_void_p(self.send32.data_ptr()),
                                            self.send32.numel(), int(self.row_items),
                                            C.c_void_p(self.sendbuf.data_ptr()), 0))
        self.tr.all_to_all(self.sendbuf, self.halo.send_counts, self.recvbuf, self.halo.recv_counts,
                           self.elems)
        p._check(p.lib.pyqed_heom_halo_pack(p._h, array_id, C.c_void_p(self.need32.data_ptr()),
                                            self.need32.numel(), int(self.row_items),
                                            C.c_void_p(self.recvbuf.data_ptr()), 1))

    def halo_bytes_per_stage(self):
        if self.native:
            return self.need64.numel() * self.elems * 8
        return self.need32.numel() * self.elems * 8

    # -- propagation ---------------------------------------------------------
    def set_state(self, rho0):
        rho0 = np.asarray(rho0, dtype=C128).reshape(1, self.n, self.n)
        if self.row_items and not np.array_equal(rho0[0], rho0[0].conj().T):
            raise ValueError("row halos assume Hermitian ADOs; rho0 is not Hermitian")
        if self.native:
            from .._cabi import _dptr
            r = np.ascontiguousarray(rho0[0])
            self.plan._check(self.plan.lib.pyqed_heom_shard_set_state(self.plan._h, _dptr(r.view(np.float64))))
            self._host_barrier()   # peers store into this rank's pools: nobody starts before all are zeroed
            return
        self.plan.set_state(rho0)
        if self.symm is not None:
            # peers store into this rank's arrays: nobody may start pushing before
            # every rank has finished (re)initialising its state
            self.symm.barrier(channel=0)

    def propagate(self, dt, nt, traj=None, fsys=None, fcoup=None):
        """RK4 for ``nt`` steps; ``traj`` (torch complex128 [1, nt+1, N, N]) is
        filled on the rank that owns the system density matrix; ``fsys`` /
        ``fcoup``: pulse tables [nt, 3] as in ``Plan.propagate``."""
        import ctypes as C
        p = self.plan
        dp = C.POINTER(C.c_double)
        if self.native:
            if fsys is not None or fcoup is not None:
                raise NotImplementedError("the rank-local sharded layout propagates time-independent "
                                          "operators only: construct ShardedDEOM(native=False) for pulses")
            return self._native_propagate(dt, nt, traj)

        def field(f):
            if f is None:
                return None, None
            f = np.ascontiguousarray(np.broadcast_to(np.asarray(f, dtype=np.float64), (1, nt, 3)))
            return f, f.ctypes.data_as(dp)
        fs, fsp = field(fsys)
        fc, fcp = field(fcoup)
        tp = None if traj is None else C.c_void_p(traj.data_ptr())
        p._check(p.lib.pyqed_heom_propagate_begin(p._h, float(dt), int(nt), fsp, fcp, tp))
        for i in range(nt):
            for st in range(4):
                p._check(p.lib.pyqed_heom_propagate_stage(p._h, i, st))
                self.exchange((1, 2, 3, 0)[st] if p.info("rk_scheme") else STAGE_OUTPUT_ARRAY[st])

    def run(self, rho0, dt, nt, pulse_system_func=None, pulse_coupling_func=None):
        """Returns ``(t_save, rho_sys[nt+1, N, N])`` on every rank."""
        from .deom import sample_pulse
        self.set_state(rho0)
        dev = self.plan._tables.device
        traj = torch.zeros((1, nt + 1, self.n, self.n), dtype=torch.complex128, device=dev)
        self.propagate(dt, nt, traj, sample_pulse(pulse_system_func, dt, nt),
                       sample_pulse(pulse_coupling_func, dt, nt))
        flat = torch.view_as_real(traj).contiguous()
        self.tr.broadcast(flat, self.owner_of_sys)
        t_save = np.arange(nt + 1, dtype=np.float64) * dt
        return t_save, torch.view_as_complex(flat)[0].cpu().numpy()

    def gather_ados(self):
        """All ADOs in reference id order on every rank (test helper: each rank
        contributes the slots it owns)."""
        p = self.plan
        if self.native:
            from .._cabi import _dptr
            import ctypes as C
            n_own = self.hi - self.lo
            own = np.zeros((max(n_own, 1), self.n, self.n), dtype=C128)
            ids = np.zeros(max(n_own, 1), dtype=np.int32)
            p._check(p.lib.pyqed_heom_shard_get_owned(p._h, _dptr(own.view(np.float64)),
                                                      ids.ctypes.data_as(C.POINTER(C.c_int32))))
            full = np.zeros((self.nmax, self.n, self.n), dtype=C128)
            full[ids[:n_own]] = own[:n_own]
            t = torch.from_numpy(np.ascontiguousarray(full).view(np.float64))
            if self.tr.nccl:
                t = t.cuda()
            self.tr.allreduce_sum(t)
            return t.cpu().numpy().view(C128).reshape(self.nmax, self.n, self.n)
        full = p.get_ados()[0]
        off = p.info("off_id_of_slot")
        id_of_slot = p._tables[off:off + 4 * self.nmax].view(torch.int32).cpu().numpy()
        owned = np.zeros(self.nmax, dtype=bool)
        owned[id_of_slot[self.lo:self.hi]] = True
        full = np.where(owned[:, None, None], full, 0)
        t = torch.from_numpy(np.ascontiguousarray(full).view(np.float64))
        if self.tr.nccl:
            t = t.cuda()
        self.tr.allreduce_sum(t)
        return t.cpu().numpy().view(C128).reshape(self.nmax, self.n, self.n)


def _lex_rank(keys, lmax):
    """Lexicographic storage slot of every multi-index (``rank_lex`` in
    ``csrc/heom_core.cuh``), vectorised on the host."""
    from math import comb
    nmax, K = keys.shape
    side = K + lmax + 2
    tab = np.array([[comb(a, b) if 0 <= b <= a else 0 for b in range(side)] for a in range(side)],
                   dtype=np.int64)
    r = np.zeros(nmax, dtype=np.int64)
    b = np.full(nmax, lmax, dtype=np.int64)
    for i in range(K):
        d = K - 1 - i
        r += tab[b + d + 1, d + 1] - tab[b - keys[:, i] + d + 1, d + 1]
        b -= keys[:, i]
    return r
