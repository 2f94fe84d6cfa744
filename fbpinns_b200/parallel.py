"""
Multi-GPU sharding of the FBPINN step: one process per GPU, subdomains partitioned into contiguous
blocks of the global subdomain index (= slabs of grid rows along dimension 0, fbpinns/decompositions.py:163).

  * network parameters, their gradients and Adam state are purely local to the owning rank — no weight all-reduce;
  * every collocation point is OWNED by the rank holding the lowest-index subdomain that contains it; the owner
    evaluates the quotient rule, the constraining operator and the loss for it;
  * a point that also lies inside subdomains of another rank (the overlap strip between two slabs) receives that
    rank's partial numerator sums in the forward pass and sends back the cotangent of the row in the reverse pass:
    ONE all_to_all_single per direction per constraint (NCCL over NVLink/NVSwitch; gloo in the CPU tests), the
    denominators D are exchanged once per active-set change;
  * the scalar loss and the gradients of the problem's own trainables are all-reduced.

Two ways to evaluate the loss (fbpinns/trainers.py:249-267 on the whole point set) on sharded points:
  * "weighted" (single-constraint problems without trainables of their own, e.g. cfg 3-5): every rank evaluates
    `loss_fn` on the points it owns; `loss_fn` must be a mean over points, so the global loss is sum_r (n_owned_r / n) L_r
    and rank r scales its cotangents by n_owned_r / n.  No extra exchange.
  * "replicated" (several constraints and / or problem trainables, e.g. cfg 1-2): the owners' ujs rows are summed into
    the full (n_ic, V) array on every rank (one all-reduce of disjoint rows per constraint: exact), every rank evaluates
    the SAME `loss_fn` on the full arrays and takes the rows it owns from the cotangent.  The loss, the gradients of the
    problem's trainables and hence their Adam updates are then bit-identical on all ranks — no weights, no assumption
    on the form of `loss_fn`, no empty-constraint special case.
"""

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import ptr, stream_ptr, check
from .engine import Plan, DeviceTakes, ConstraintEvaluator, gather_rows, nonzero_i32


class Shard:
    def __init__(self, rank, world, group=None, overlap_halo=None):
        self.rank, self.world, self.group = int(rank), int(world), group
        # split-phase halo exchange overlapped with the interior work items (boundary items launched first).  Measured
        # on 2 x B200 (cfg 5): 187.5 vs 191.1 steps/s with the plain exchange — the extra partial-wave tail of two
        # launches per kernel costs more than the ~30 us of all_to_all latency it hides — so it is OFF by default.
        if overlap_halo is None:
            overlap_halo = os.environ.get("FBP_HALO_OVERLAP", "0") == "1"
        self.overlap_halo = bool(overlap_halo)

        self.bounds = None          # optional explicit block boundaries (world + 1 ascending subdomain indices)
        # halo transport: "peer" = direct stores into the owner's buffer over NVLink peer memory (fbp_halo_*), "nccl" =
        # all_to_all_single; default: peer on CUDA when symmetric memory is available (FBP_HALO=nccl forces the collective)
        self.halo_transport = os.environ.get("FBP_HALO", "peer")

    def block(self, m, j=None):
        "contiguous block [lo, hi) of global subdomain indices owned by rank j"
        j = self.rank if j is None else j
        if self.bounds is not None:
            assert len(self.bounds) == self.world + 1 and self.bounds[-1] == m
            return int(self.bounds[j]), int(self.bounds[j + 1])
        return (j * m) // self.world, ((j + 1) * m) // self.world

    def balance(self, pairs_per_subdomain):
        """Block boundaries that balance the PAIR count (= the kernels' work) instead of the subdomain count: edge slabs of
        a regular decomposition hold ~8 % fewer pairs than interior ones.  Computed once from the full point set so that
        parameter / Adam-state ownership never moves; every rank computes the same boundaries."""
        c = np.concatenate([[0], np.cumsum(np.asarray(pairs_per_subdomain, dtype=np.int64))])
        m, tot = len(c) - 1, int(c[-1])
        b = [0]
        for j in range(1, self.world):
            k = int(np.searchsorted(c, tot * j / self.world, side="left"))
            # the boundary whose cumulative count is closest to the target, at least one subdomain per rank
            if k > 0 and abs(c[k - 1] - tot * j / self.world) <= abs(c[min(k, m)] - tot * j / self.world):
                k -= 1
            b.append(min(max(k, b[-1] + 1), m - (self.world - j)))
        self.bounds = np.asarray(b + [m], dtype=np.int64)
        return self.bounds


def shard_trainer(trainer, rank, world, group=None):
    trainer.shard = Shard(rank, world, group)
    return trainer


# --------------------------------------------------------------------------------------------------- halo bookkeeping (pure index logic)

def build_halo_lists(inside, rank):
    """inside: (world, n) bool — inside[j, p] = point p lies in >= 1 subdomain of rank j (of the current all_ims).
    Returns dict with
      local_ips   ascending indices of the points this rank evaluates (inside[rank])
      owner       (n,) owning rank of every point (lowest rank containing it, -1 if none)
      owned_local bool over local points: this rank owns the point
      send[j]     local indices whose owner is j (this rank sends its partial sums for them to j)
      recv[j]     local indices owned here that rank j also touches (partials arrive from j)
    send[j] on rank r and recv[r] on rank j enumerate the same points in the same (ascending global) order."""
    inside = np.asarray(inside, dtype=bool)
    world, n = inside.shape
    covered = inside.any(0)
    owner = np.where(covered, inside.argmax(0), -1)
    local_ips = np.nonzero(inside[rank])[0]
    own_l = owner[local_ips] == rank
    send, recv = {}, {}
    for j in range(world):
        if j == rank:
            continue
        send[j] = np.nonzero(owner[local_ips] == j)[0]
        recv[j] = np.nonzero(own_l & inside[j][local_ips])[0]
    return dict(local_ips=local_ips, owner=owner, owned_local=own_l, send=send, recv=recv)


class HaloExchange:
    """Row exchange for one constraint: forward adds the partial sums of shared rows into their owner,
    reverse returns the owner's row cotangents to the sharers."""

    def __init__(self, halo, shard, device):
        self.shard = shard
        w, r = shard.world, shard.rank
        tl = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int64), dtype=torch.long, device=device)
        self.send_counts = [0 if j == r else len(halo["send"][j]) for j in range(w)]
        self.recv_counts = [0 if j == r else len(halo["recv"][j]) for j in range(w)]
        self.send_idx = tl(np.concatenate([halo["send"][j] for j in range(w) if j != r]) if w > 1 else [])
        self.recv_idx = tl(np.concatenate([halo["recv"][j] for j in range(w) if j != r]) if w > 1 else [])
        self.active = (sum(self.send_counts) + sum(self.recv_counts)) > 0

    def _a2a(self, out, inp, out_splits, in_splits, async_op=False):
        return dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits,
                                      group=self.shard.group, async_op=async_op)

    # split-phase variants: start() enqueues the exchange on NCCL's stream and returns immediately, so kernels launched
    # afterwards on the compute stream overlap it; finish() makes the compute stream wait and applies the result
    def forward_start(self, rows):
        inp = rows.index_select(0, self.send_idx).contiguous()
        out = torch.empty((int(self.recv_idx.numel()), rows.shape[1]), dtype=rows.dtype, device=rows.device)
        work = self._a2a(out, inp, list(self.recv_counts), list(self.send_counts), async_op=True)
        return work, inp, out

    def forward_finish(self, rows, pending):
        work, _, out = pending
        work.wait()
        off = 0
        for c in self.recv_counts:
            if c:
                rows.index_add_(0, self.recv_idx[off:off + c], out[off:off + c])
                off += c
        return rows

    def backward_start(self, rows):
        inp = rows.index_select(0, self.recv_idx).contiguous()
        out = torch.empty((int(self.send_idx.numel()), rows.shape[1]), dtype=rows.dtype, device=rows.device)
        work = self._a2a(out, inp, list(self.send_counts), list(self.recv_counts), async_op=True)
        return work, inp, out

    def backward_finish(self, rows, pending):
        work, _, out = pending
        work.wait()
        rows.index_copy_(0, self.send_idx, out)
        return rows

    def forward_add(self, rows):
        "rows (q, V): adds into owned shared rows the partial sums the other ranks computed for them (in rank order)"
        V = rows.shape[1]
        inp = rows.index_select(0, self.send_idx).contiguous()
        out = torch.empty((int(self.recv_idx.numel()), V), dtype=rows.dtype, device=rows.device)
        self._a2a(out, inp, [c for c in self.recv_counts], [c for c in self.send_counts])
        # a row can be shared with more than one rank: accumulate peer by peer (deterministic order)
        off = 0
        for c in self.recv_counts:
            if c:
                rows.index_add_(0, self.recv_idx[off:off + c], out[off:off + c])
                off += c
        return rows

    def backward_return(self, rows):
        "rows (q, V): overwrites the rows owned elsewhere with the values their owners hold"
        V = rows.shape[1]
        inp = rows.index_select(0, self.recv_idx).contiguous()
        out = torch.empty((int(self.send_idx.numel()), V), dtype=rows.dtype, device=rows.device)
        self._a2a(out, inp, [c for c in self.send_counts], [c for c in self.recv_counts])
        rows.index_copy_(0, self.send_idx, out)
        return rows


def peer_halo_layout(halo, cnt, rank, rows_per_block=256):
    """Pure index logic of the peer-memory exchange for one rank (tested on the CPU by simulating every rank,
    tests/test_parallel_cpu.py).  halo = build_halo_lists(...)[rank]; cnt[a][b] = rows rank a sends to owner b in the
    forward direction (the same matrix on every rank).  Receive region of rank b, in rows: [forward: senders in rank order]
    [reverse: owners in rank order].  Returns
      fwd_send / bwd_send   local rows to send (forward: rows owned elsewhere; reverse: owned rows others share), peers in rank order
      rows_fwd_dst[j] / rows_bwd_dst[j]   row offset in peer j's region such that position r of the send list lands at row offset + r
      fwd_blocks / bwd_blocks             (peer, first position, end position, blocks of that peer) per CTA of the push kernel
      fwd_tgt, fwd_ptr, fwd_pos           owned shared rows and, CSR, the forward-region rows to add into each (senders in rank order)
      bwd_ptr, bwd_pos                    reverse pull: row t of the reverse region is the value of send-list row t
      fwd_mask / bwd_mask                 peers data arrives from;  my_bwd_off_rows  start of this rank's reverse region"""
    cnt = np.asarray(cnt, dtype=np.int64)
    w, r = cnt.shape[0], int(rank)
    others = [j for j in range(w) if j != r]
    fwd_total = cnt.sum(0)                                # rows rank a receives in the forward direction
    cat = lambda parts: np.concatenate(parts).astype(np.int64) if len(parts) else np.zeros(0, dtype=np.int64)
    fwd_send = cat([np.asarray(halo["send"][j]) for j in others])
    bwd_send = cat([np.asarray(halo["recv"][j]) for j in others])
    fs_off = np.concatenate([[0], np.cumsum([len(halo["send"][j]) for j in others])]).astype(np.int64)
    bs_off = np.concatenate([[0], np.cumsum([len(halo["recv"][j]) for j in others])]).astype(np.int64)
    rows_fwd_dst, rows_bwd_dst = np.zeros(w, dtype=np.int64), np.zeros(w, dtype=np.int64)
    fblk, bblk = [], []
    for k, j in enumerate(others):
        rows_fwd_dst[j] = int(cnt[:r, j].sum()) - int(fs_off[k])                       # after the rows of lower-ranked senders
        rows_bwd_dst[j] = int(fwd_total[j]) + int(cnt[j, :r].sum()) - int(bs_off[k])   # after j's forward region and lower owners
        for lst, off, blk in ((halo["send"][j], fs_off, fblk), (halo["recv"][j], bs_off, bblk)):
            n_ = len(lst)
            nb = -(-n_ // rows_per_block)
            for b_ in range(nb):
                blk.append((j, int(off[k]) + b_ * rows_per_block, int(off[k]) + min(n_, (b_ + 1) * rows_per_block), nb))
    pos_of, base = {}, 0
    for j in others:
        for k_, row in enumerate(halo["recv"][j]):
            pos_of.setdefault(int(row), []).append(base + k_)
        base += len(halo["recv"][j])
    tg = sorted(pos_of)
    return dict(fwd_send=fwd_send, bwd_send=bwd_send, rows_fwd_dst=rows_fwd_dst, rows_bwd_dst=rows_bwd_dst,
                fwd_blocks=np.asarray(fblk, dtype=np.int64).reshape(-1, 4), bwd_blocks=np.asarray(bblk, dtype=np.int64).reshape(-1, 4),
                fwd_tgt=np.asarray(tg, dtype=np.int64), fwd_ptr=np.concatenate([[0], np.cumsum([len(pos_of[t]) for t in tg])]).astype(np.int64),
                fwd_pos=np.asarray([p for t in tg for p in pos_of[t]], dtype=np.int64),
                bwd_ptr=np.arange(len(fwd_send) + 1), bwd_pos=np.arange(len(fwd_send)),
                fwd_mask=sum(1 << j for j in others if len(halo["recv"][j])), bwd_mask=sum(1 << j for j in others if len(halo["send"][j])),
                my_bwd_off_rows=int(fwd_total[r]), region_rows=int(fwd_total[r] + cnt[r].sum()))


class PeerHaloExchange:
    """HaloExchange over NVLink peer memory (C ABI fbp_halo_push / fbp_halo_pull, csrc/fbp_halo.cu): the sender stores its
    rows straight into the owner's receive buffer (torch.distributed._symmetric_memory allocation), publishes a flag, and
    the owner sums them into its rows in peer order; the reverse pass returns the cotangents the same way.  Two kernels
    per exchange instead of index_select + all_to_all_single + up to 7 index_add_ launches, no NCCL latency."""

    ROWS_PER_BLOCK = 256          # rows per CTA of the push kernel (cfg 5, 2 ranks: 32 K shared rows -> 128 CTAs)

    def __init__(self, halo, shard, device, row_floats_max):
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import HaloPeers
        self.shard = shard
        w, r = shard.world, shard.rank
        if w > _lib.FBP_HALO_MAX_WORLD:
            raise _lib.FbpError(f"peer halo exchange supports up to {_lib.FBP_HALO_MAX_WORLD} ranks")
        group = shard.group if shard.group is not None else dist.group.WORLD
        mine = torch.tensor([0 if j == r else len(halo["send"][j]) for j in range(w)], dtype=torch.int64, device=device)
        allc = [torch.zeros_like(mine) for _ in range(w)]
        dist.all_gather(allc, mine, group=group)
        cnt = torch.stack(allc).cpu().numpy()                 # cnt[a][b] = rows a sends to owner b (forward)
        assert all(cnt[j][r] == len(halo["recv"][j]) for j in range(w) if j != r)
        fwd_total = cnt.sum(0)                                # rows rank a receives in the forward direction
        bwd_total = cnt.sum(1)                                # ... in the reverse direction (as a sharer)
        cap_rows = int((fwd_total + bwd_total).max())
        self.flag_floats = 64                                 # 256 bytes: int32 flags[4][8]
        buf = symm_mem.empty(self.flag_floats + max(cap_rows, 1) * int(row_floats_max), dtype=torch.float32, device=device)
        buf.zero_()
        torch.cuda.synchronize()
        self.hdl = symm_mem.rendezvous(buf, group)
        self.buf = buf
        self.peers = HaloPeers()
        for j in range(w):
            self.peers.flags[j] = int(self.hdl.buffer_ptrs[j])
            self.peers.data[j] = int(self.hdl.buffer_ptrs[j]) + 4 * self.flag_floats
        i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32), dtype=torch.int32, device=device)
        lay = peer_halo_layout(halo, cnt, r, self.ROWS_PER_BLOCK)
        self.fwd_send, self.bwd_send = i32(lay["fwd_send"]), i32(lay["bwd_send"])
        self.rows_fwd_dst, self.rows_bwd_dst = lay["rows_fwd_dst"], lay["rows_bwd_dst"]
        self.fwd_blocks, self.n_fwd_blocks = i32(lay["fwd_blocks"]), len(lay["fwd_blocks"])
        self.bwd_blocks, self.n_bwd_blocks = i32(lay["bwd_blocks"]), len(lay["bwd_blocks"])
        self.fwd_tgt, self.fwd_ptr, self.fwd_pos, self.fwd_mask = i32(lay["fwd_tgt"]), i32(lay["fwd_ptr"]), i32(lay["fwd_pos"]), lay["fwd_mask"]
        self.bwd_tgt, self.bwd_ptr, self.bwd_pos, self.bwd_mask = self.fwd_send, i32(lay["bwd_ptr"]), i32(lay["bwd_pos"]), lay["bwd_mask"]
        self.my_bwd_off_rows = lay["my_bwd_off_rows"]
        self.epoch = torch.zeros(2, dtype=torch.int32, device=device)
        self.ticket = torch.zeros(_lib.FBP_HALO_MAX_WORLD, dtype=torch.int32, device=device)
        self.done = torch.zeros(1, dtype=torch.int32, device=device)
        self._dst_cache = {}
        self.active = True
        dist.barrier(group=group)               # every rank has zeroed its flags before anyone pushes

    def _dst(self, rows_off, V):
        key = (id(rows_off), V)
        if key not in self._dst_cache:
            self._dst_cache[key] = torch.as_tensor(rows_off * V, dtype=torch.int64, device=self.epoch.device)
        return self._dst_cache[key]

    def _exchange(self, rows, direction):
        lib = _lib.load()
        V = int(rows.shape[1])
        r, w = self.shard.rank, self.shard.world
        fwd = direction == 0
        send, blocks, nb = (self.fwd_send, self.fwd_blocks, self.n_fwd_blocks) if fwd else (self.bwd_send, self.bwd_blocks, self.n_bwd_blocks)
        check(lib.fbp_halo_push(ptr(rows), V, ptr(send), ptr(blocks), nb, C.byref(self.peers),
                                C.c_void_p(self._dst(self.rows_fwd_dst if fwd else self.rows_bwd_dst, V).data_ptr()), r, direction,
                                ptr(self.epoch), ptr(self.ticket), stream_ptr()), "fbp_halo_push")
        tgt, sp, pos, mask = (self.fwd_tgt, self.fwd_ptr, self.fwd_pos, self.fwd_mask) if fwd else (self.bwd_tgt, self.bwd_ptr, self.bwd_pos, self.bwd_mask)
        check(lib.fbp_halo_pull(ptr(rows), V, ptr(tgt), ptr(sp), ptr(pos), int(tgt.numel()), C.byref(self.peers),
                                0 if fwd else self.my_bwd_off_rows * V, mask, r, w, direction, 0 if fwd else 1,
                                ptr(self.epoch), ptr(self.done), stream_ptr()), "fbp_halo_pull")
        return rows

    def forward_add(self, rows):
        "rows (q, V) contiguous: adds into owned shared rows the partial sums the other ranks computed for them"
        return self._exchange(rows, 0)

    def backward_return(self, rows):
        "rows (q, V) contiguous: overwrites the rows owned elsewhere with the values their owners hold"
        return self._exchange(rows, 1)


def make_halo_exchange(halo, shard, device, row_floats_max):
    "peer-memory transport on CUDA (unless FBP_HALO=nccl), the collective otherwise (gloo in the CPU tests)"
    if torch.device(device).type == "cuda" and shard.world > 1 and shard.halo_transport == "peer":
        return PeerHaloExchange(halo, shard, device, row_floats_max)
    return HaloExchange(halo, shard, device)


# --------------------------------------------------------------------------------------------------- sharded evaluator

class ShardedEvaluator:
    "ConstraintEvaluator over this rank's subdomains and points + the halo exchange with the other ranks"

    def __init__(self, ev: ConstraintEvaluator, halo, shard):
        self.ev, self.shard = ev, shard
        dev = ev.x.device
        self.halo = make_halo_exchange(halo, shard, dev, max(ev.V, ev.plan.jet.C) * int(ev.takes.npou))
        self.owned_idx = torch.as_tensor(np.nonzero(halo["owned_local"])[0], dtype=torch.long, device=dev)
        inv = -np.ones(len(halo["owned_local"]), dtype=np.int32)
        inv[halo["owned_local"]] = np.arange(int(halo["owned_local"].sum()), dtype=np.int32)
        self.owned_inv = torch.as_tensor(inv, dtype=torch.int32, device=dev)
        t = ev.takes
        self.npou = int(t.npou)
        self.nsum = torch.empty((max(t.q, 1), ev.V), dtype=torch.float32, device=dev)[:t.q]
        if self.npou == 1:
            if t.q != t.n:
                raise _lib.FbpError("sharded evaluation: a single partition of unity must give one row per local point")
            # denominators: local partial -> total on the owners (once per active-set change)
            self.halo.forward_add(ev.dsum[:t.q])          # collective: every rank calls it, even with no rows
        else:
            self._setup_multilevel()
        self._split_work_list()

    # ---- multilevel decompositions (npou > 1): rows are (point, level) pairs and the owner of a point may hold no row for a
    # level its own subdomains do not cover there.  Rows are therefore exchanged through a DENSE per-point layout
    # (n_local, npou, V) (zeros where this rank has no row), and the owner applies the per-level quotient rule and the average
    # over levels (fbpinns/trainers.py:160-170) on that layout in torch (differentiable; this path is about function, the
    # single-level path is the tuned one).
    def _setup_multilevel(self):
        ev, t = self.ev, self.ev.takes
        dev = ev.x.device
        m_take, n_take, p_take, np_take, npou = t.reference_arrays()
        pou_ids = ev.decomp.pou_host[t.sub_ids.cpu().numpy()]                    # level id of every local subdomain
        levels = np.unique(ev.decomp.pou_host)
        assert len(levels) == npou
        row_level = np.zeros(t.q, dtype=np.int64)
        row_level[p_take] = np.searchsorted(levels, pou_ids[m_take])
        self.dense_idx = torch.as_tensor(np_take.astype(np.int64) * npou + row_level, dtype=torch.long, device=dev)
        C = ev.plan.jet.C
        dd = torch.zeros((t.n * npou, C), dtype=torch.float32, device=dev)
        dd.index_copy_(0, self.dense_idx, ev.dsum[:t.q])
        dd = dd.view(t.n, npou * C)
        self.halo.forward_add(dd)
        self.dsum_dense_owned = dd.index_select(0, self.owned_idx).view(-1, npou, C).contiguous()
        self.overlap = False

    def quotient(self, nd):
        "owned dense numerator jets (n_owned, npou * V) -> ujets (n_owned, V): per-level quotient rule, average over levels"
        jet = self.ev.plan.jet
        n, L, C, ud = nd.shape[0], self.npou, jet.C, jet.ud
        N = nd.view(n, L, C, ud)
        D = self.dsum_dense_owned.unsqueeze(-1)                                  # (n, L, C, 1)
        q = [None] * C
        d0 = D[:, :, 0]
        for c, path in enumerate(jet.comps):
            if len(path) == 0:
                q[c] = N[:, :, c] / d0
        c0 = jet.index[()]
        for c, path in enumerate(jet.comps):
            if len(path) == 1:
                q[c] = (N[:, :, c] - q[c0] * D[:, :, c]) / d0
        for c, path in enumerate(jet.comps):
            if len(path) == 2:
                ck, cl = jet.index[(path[0],)], jet.index[(path[1],)]
                q[c] = (N[:, :, c] - q[ck] * D[:, :, cl] - q[cl] * D[:, :, ck] - q[c0] * D[:, :, c]) / d0
        return torch.stack(q, dim=2).mean(dim=1).reshape(n, C * ud)

    def _split_work_list(self):
        """Boundary work items = items of subdomains that hold at least one row shared with another rank.  They are
        launched first in the forward pass (their row sums feed the halo exchange, which then overlaps the interior
        items) and last in the reverse pass (they need the cotangents that come back).  Tiled plans only."""
        ev, t = self.ev, self.ev.takes
        self.overlap = (self.shard.overlap_halo and bool(ev.plan.is_fast) and t.s > 0 and self.shard.world > 1
                        and isinstance(self.halo, HaloExchange))       # the split-phase variant exists for the collective only
        if not self.overlap:
            return
        dev = ev.x.device
        halo_row = torch.zeros(max(t.q, 1), dtype=torch.bool, device=dev)
        halo_row[self.halo.send_idx] = True
        halo_row[self.halo.recv_idx] = True
        bsub = torch.zeros(t.m_all, dtype=torch.bool, device=dev)
        bsub[t.spair_sub.long()[halo_row[t.spair_row.long()]]] = True
        bsub = bsub.cpu().numpy()
        items = t.items_host
        is_b = bsub[items[:, 0]]
        of, ob = t.item_order_fwd.cpu().numpy(), t.item_order_bwd.cpu().numpy()
        mk = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32), dtype=torch.int32, device=dev)
        self._orders = [mk(of[is_b[of]]), mk(of[~is_b[of]]), mk(ob[~is_b[ob]]), mk(ob[is_b[ob]])]
        base = t.view()

        def view(order, fwd):
            v = type(base).from_buffer_copy(base)
            if fwd:
                v.d_item_order_fwd, v.n_items = (order.data_ptr() if order.numel() else None), int(order.numel())
            else:
                v.d_item_order_bwd, v.n_items_active = (order.data_ptr() if order.numel() else None), int(order.numel())
            return v
        self.v_fwd_boundary, self.v_fwd_interior = view(self._orders[0], True), view(self._orders[1], True)
        self.v_bwd_interior, self.v_bwd_boundary = view(self._orders[2], False), view(self._orders[3], False)

    def forward(self, params):
        lib = _lib.load()
        ev = self.ev
        tv = ev.takes.view()

        def fwd(view):
            check(lib.fbp_forward(ev.plan.handle, C.byref(view), ptr(ev.x), ptr(params), ptr(ev.decomp.sub_static),
                                  ptr(ev.pair_out), ptr(ev.scratch), ev.scratch_floats, ptr(ev.cache), stream_ptr()),
                  "fbp_forward")

        def row_sums():
            check(lib.fbp_row_sums(ev.plan.handle, C.byref(tv), ptr(ev.pair_out), ptr(self.nsum), stream_ptr()),
                  "fbp_row_sums")
        if self.overlap:
            fwd(self.v_fwd_boundary)                      # subdomains touching the slab interfaces first
            row_sums()                                    # shared rows are complete now (interior rows are not: unused)
            pending = self.halo.forward_start(self.nsum)  # exchange runs on NCCL's stream ...
            fwd(self.v_fwd_interior)                      # ... while the interior subdomains are evaluated
            row_sums()
            self.halo.forward_finish(self.nsum, pending)
        elif self.npou > 1:
            fwd(tv)
            row_sums()
            t = ev.takes
            dense = torch.zeros((t.n * self.npou, ev.V), dtype=torch.float32, device=ev.x.device)
            dense.index_copy_(0, self.dense_idx, self.nsum)
            dense = dense.view(t.n, self.npou * ev.V)
            self.halo.forward_add(dense)
            return dense.index_select(0, self.owned_idx)          # (n_owned, npou * V): quotient() follows outside
        else:
            fwd(tv)
            row_sums()
            self.halo.forward_add(self.nsum)
        # quotient rule (+ fused affine constraining) for the OWNED points only, written compactly
        ujets = torch.empty((int(self.owned_idx.numel()), ev.V), dtype=torch.float32, device=ev.x.device)
        check(lib.fbp_reduce_rows_forward(ev.plan.handle, C.byref(tv), ptr(self.nsum), ptr(ev.dsum), ptr(ev.affine),
                                          ptr(self.owned_inv), ptr(ujets), stream_ptr()), "fbp_reduce_rows_forward")
        return ujets

    def backward(self, ujets_bar_owned, params, grads, weight=1.0):
        lib = _lib.load()
        ev = self.ev
        tv = ev.takes.view()
        if self.npou > 1:
            # cotangent of the dense numerators of the owned points -> all local points -> back to the sharers -> rows
            t = ev.takes
            W = self.npou * ev.V
            gd = torch.empty((t.n, W), dtype=torch.float32, device=ev.x.device)
            src = ujets_bar_owned.contiguous().float()
            check(lib.fbp_scatter_rows(ptr(src), ptr(self.owned_inv), t.n, W, float(weight), ptr(gd), stream_ptr()),
                  "fbp_scatter_rows")
            self.halo.backward_return(gd)
            ev.grow[:t.q].copy_(gd.view(t.n * self.npou, ev.V).index_select(0, self.dense_idx))
            check(lib.fbp_backward(ev.plan.handle, C.byref(tv), ptr(ev.x), ptr(params), ptr(ev.decomp.sub_static),
                                   ptr(ev.grow), ptr(grads), 1, ptr(ev.gpart), ptr(ev.scratch), ev.scratch_floats,
                                   ptr(ev.cache), stream_ptr()), "fbp_backward")
            return
        # cotangent of every local point: the owned ones scaled by the ownership weight, zero for the others (one launch)
        ub = torch.empty((ev.takes.n, ev.V), dtype=torch.float32, device=ev.x.device)
        src = ujets_bar_owned.contiguous().float()
        check(lib.fbp_scatter_rows(ptr(src), ptr(self.owned_inv), ev.takes.n, ev.V, float(weight), ptr(ub), stream_ptr()),
              "fbp_scatter_rows")
        check(lib.fbp_reduce_backward(ev.plan.handle, C.byref(tv), ptr(ub), ptr(ev.dsum), ptr(ev.affine), ptr(ev.grow),
                                      stream_ptr()), "fbp_reduce_backward")
        def bwd(view, flags):
            check(lib.fbp_backward(ev.plan.handle, C.byref(view), ptr(ev.x), ptr(params), ptr(ev.decomp.sub_static),
                                   ptr(ev.grow), ptr(grads), flags, ptr(ev.gpart), ptr(ev.scratch), ev.scratch_floats,
                                   ptr(ev.cache), stream_ptr()), "fbp_backward")
        rows = ev.grow[:ev.takes.q]
        if self.overlap:
            pending = self.halo.backward_start(rows)      # owners' cotangents of the shared rows travel back ...
            bwd(self.v_bwd_interior, 2)                   # ... while the interior subdomains run (kernels only)
            self.halo.backward_finish(rows, pending)
            bwd(self.v_bwd_boundary, 2)
            bwd(tv, 4 | 1)                                # sum every item's partial gradients, add into grads
        else:
            self.halo.backward_return(rows)
            bwd(tv, 1)


class _ShardedSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tape_hook, sev, params, grads, weight):
        ctx.sev, ctx.params, ctx.grads, ctx.weight = sev, params, grads, weight
        return sev.forward(params)

    @staticmethod
    def backward(ctx, ubar):
        ctx.sev.backward(ubar, ctx.params, ctx.grads, ctx.weight)
        return None, None, None, None, None


def sharded_sum(sev, params, grads, tape_hook, weight):
    out = _ShardedSum.apply(tape_hook, sev, params, grads, weight)
    return sev.quotient(out) if sev.npou > 1 else out


class _ReplicateRows(torch.autograd.Function):
    """(n_owned, V) rows of the owners -> the full (n, V) array on every rank (sum of disjoint rows: exact).  Every rank
    then computes the same scalar from it, so the cotangent of the full array is the same everywhere and the reverse pass
    is a row selection without communication."""

    @staticmethod
    def forward(ctx, rows, owned_idx, n, group):
        ctx.owned_idx = owned_idx
        full = torch.zeros((n, rows.shape[1]), dtype=rows.dtype, device=rows.device)
        full.index_copy_(0, owned_idx, rows)
        dist.all_reduce(full, group=group)
        return full

    @staticmethod
    def backward(ctx, gfull):
        return gfull.index_select(0, ctx.owned_idx), None, None, None


def replicate_rows(rows, owned_idx, n, group=None):
    return _ReplicateRows.apply(rows, owned_idx, n, group)


# --------------------------------------------------------------------------------------------------- update inputs / step

def get_update_inputs_sharded(shard, active, all_params, dd, x_batch_global, constraints_global, constraint_offsets,
                              jets, layer_sizes, kernel="auto", activation="tanh", replicated=False, plans=None):
    """Sharded counterpart of trainers.get_update_inputs: the global active-set algebra is identical on every rank
    (every rank holds the full point set and the full static decomposition — both small), then takes are built for
    this rank's block of subdomains over the points inside them."""
    from .trainers import UpdateInputs, active_set_algebra
    m = dd.m
    active = np.array(active).copy()
    dev = dd.device
    ims1 = torch.as_tensor(np.arange(m, dtype=np.int32)[active == 1], dtype=torch.int32, device=dev)
    pt_count, mc1 = dd.inside_count(x_batch_global, models=ims1)
    training_ips = nonzero_i32(pt_count)
    d_stat = float(mc1.double().mean().item() ** (1 / dd.xd)) if ims1.numel() else float("nan")
    x_batch = gather_rows(x_batch_global, training_ips)
    ips_host = training_ips.cpu().numpy().astype(np.int64)
    sizes = [c_[0].shape[0] for c_ in constraints_global]
    bounds = np.searchsorted(ips_host, np.concatenate([constraint_offsets, [constraint_offsets[-1] + sizes[-1]]]))
    _, model_count = dd.inside_count(x_batch, models=None)
    active2, active_ims, fixed_ims, all_ims, _ = active_set_algebra(active, model_count.cpu().numpy())

    blocks = [shard.block(m, j) for j in range(shard.world)]
    lo, hi = blocks[shard.rank]
    a_loc = active_ims[(active_ims >= lo) & (active_ims < hi)]
    f_loc = fixed_ims[(fixed_ims >= lo) & (fixed_ims < hi)]
    all_loc = np.concatenate([a_loc, f_loc]).astype(np.int32)
    pos_loc = -np.ones(m, dtype=np.int32)
    pos_loc[all_loc] = np.arange(len(all_loc), dtype=np.int32)

    out = UpdateInputs()
    out.active, out.active_ims, out.fixed_ims, out.all_ims = active2, a_loc.astype(np.int32), f_loc.astype(np.int32), all_loc
    out.global_active_ims, out.global_all_ims = active_ims, all_ims
    out.pos_of_model, out.training_ips, out.d, out.x_batch = pos_loc, training_ips, d_stat, x_batch
    out.constraints, out.takess, out.evaluators, out.weights, out.halos = [], [], [], [], []
    out.replicated, out.owned_global, out.n_constraint = bool(replicated), [], []
    out.kernel_point_rows = []          # rows of each constraint's global arrays behind this rank's kernel points
    sorted_all = np.sort(all_ims)
    for ic in range(len(constraints_global)):
        a, b = int(bounds[ic]), int(bounds[ic + 1])
        local = (training_ips[a:b] - int(constraint_offsets[ic])).contiguous()
        con = [gather_rows(c_, local) for c_ in constraints_global[ic]]
        x_ic = con[0]
        inside = []
        for (l, h) in blocks:
            ims_j = sorted_all[(sorted_all >= l) & (sorted_all < h)].astype(np.int32)
            if len(ims_j) and x_ic.shape[0]:
                ptj, _ = dd.inside_count(x_ic, models=torch.as_tensor(ims_j, dtype=torch.int32, device=dev),
                                         want_model_count=False)
                inside.append((ptj > 0).cpu().numpy())
            else:
                inside.append(np.zeros(x_ic.shape[0], dtype=bool))
        halo = build_halo_lists(np.stack(inside), shard.rank)
        lips = torch.as_tensor(halo["local_ips"].astype(np.int32), dtype=torch.int32, device=dev)
        x_loc = gather_rows(x_ic, lips)
        out.kernel_point_rows.append(local[lips.long()].contiguous())
        plan = plans[ic] if plans is not None else Plan(layer_sizes, jets[ic], kernel=kernel, activation=activation)
        takes = DeviceTakes(dd, x_loc, pos_loc, all_loc, len(a_loc), tile_points=plan.tile_points)
        ev = ConstraintEvaluator(plan, takes, x_loc, dd)
        sev = ShardedEvaluator(ev, halo, shard)
        owned_global = torch.as_tensor(halo["local_ips"][halo["owned_local"]].astype(np.int32), dtype=torch.int32, device=dev)
        # the loss sees the owned points only ("weighted") or every rank sees all points of the constraint ("replicated")
        out.constraints.append(con if replicated else [gather_rows(c_, owned_global) for c_ in con])
        out.owned_global.append(owned_global.long())
        out.n_constraint.append(int(x_ic.shape[0]))
        out.takess.append(takes)
        out.evaluators.append(sev)
        out.halos.append(halo)
        n_ic = x_ic.shape[0]
        out.weights.append(1.0 if replicated else float(len(owned_global)) / float(max(n_ic, 1)))
    return out


class ShardedLoss:
    """The loss of a sharded step in "weighted" mode: every rank holds its share; reading the value (`item()` / `float()`)
    sums the shares over the ranks.  Reading is a COLLECTIVE: every rank must read the same steps' losses in the same order
    (the trainer's reporting and bench.py do)."""

    def __init__(self, share, group):
        self.share, self.group = share, group

    def item(self):
        t = self.share.detach().clone().reshape(1)
        dist.all_reduce(t, group=self.group)
        return t.item()

    def __float__(self):
        return float(self.item())


def make_sharded_update(base_cls):
    """UpdateStep variant whose forward goes through the sharded evaluators and which all-reduces the loss and the
    problem-parameter gradients."""

    class ShardedUpdateStep(base_cls):
        def __init__(self, shard, *a, **kw):
            super().__init__(*a, **kw)
            self.shard = shard

        def forward_loss(self):
            self._refresh_problem_views()
            cons, owned_nothing, tape = [], False, []
            for ic, (sev, con, w, aff) in enumerate(zip(self.inp.evaluators, self.inp.constraints, self.inp.weights, self.affine)):
                ujets = sharded_sum(sev, self.params, self.grads, self.hook, w)
                tape.append(ujets)
                owned_nothing = owned_nothing or (not self.inp.replicated and ujets.shape[0] == 0)
                if self.inp.replicated:
                    ujets = replicate_rows(ujets, self.inp.owned_global[ic], self.inp.n_constraint[ic], self.shard.group)
                jet = sev.ev.plan.jet
                if aff is not None:
                    ujs = jet.ujs_plain(ujets)        # the kernels already returned the constrained jets
                elif self.has_constraining:
                    ujs = jet.ujs_constrained(ujets, con[0], self.problem.constraining_fn, self.all_params)
                else:
                    ujs = jet.ujs_plain(ujets)
                cons.append(list(con) + ujs)
            if owned_nothing:
                # "weighted" mode on a rank that owns no point of a constraint (e.g. every point also lies in a subdomain of a
                # lower rank): its share of the loss is zero — a mean over an empty set would be NaN.  The zero stays attached
                # to the tape so that the reverse pass still serves the halo exchange and this rank's subdomains.
                return sum(u.sum() for u in tape) * 0.0
            return self.problem.loss_fn(self.all_params, cons)

        def _eager(self):
            if self.inp.replicated:
                # every rank holds the global loss and identical problem-parameter gradients: nothing to reduce
                return base_cls._eager(self)
            self.grads.zero_()
            self.hook.grad = None
            loss = self.forward_loss()
            loss.backward()
            with torch.no_grad():
                self.adam.step(self.params, self.grads, self.active_ims_dev, None, None)
                # this rank's share of the global loss (sum over ranks of n_owned / n times the local mean); the sum over the
                # ranks is only formed when the value is read (ShardedLoss.item): nothing in the step depends on it
                self.loss_out.copy_(loss.detach() * self.inp.weights[0])
            return self.loss_out

        def __call__(self):
            out = base_cls.__call__(self)
            return out if self.inp.replicated else ShardedLoss(out, self.shard.group)

    return ShardedUpdateStep
