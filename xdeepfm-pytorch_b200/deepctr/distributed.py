"""Single-node multi-GPU training: one process per GPU (torchrun), NCCL for the plumbing, NVLink peer memory for the tables.

Replaces the reference's single-process `torch.nn.DataParallel` (deepctr/models/basemodel.py:206-209), which replicates the
whole model -- every embedding table included -- on every GPU each step.  Here

  * dense parameters (CIN, DNN, heads, attention, first-order dense weight) are data-parallel: the flat gradient buffer of the
    fused optimizer is summed with ONE NCCL all-reduce per step.  Losses are `reduction='sum'` (basemodel.py:254), so gradients
    are SUMMED (not averaged) and the L2 term is applied once, after the reduction -- an N-GPU step on N local batches equals
    a 1-GPU step on their concatenation.
  * embedding tables and the [V,1] first-order tables are ROW-SHARDED: global row r of a table lives on rank r % G at local row
    r // G.  Lookups read remote rows directly through peer-mapped memory inside the gather kernel; gradients are reduced per
    rank, left in an exported exchange buffer and pulled by the owning rank, which merges them deterministically and runs the
    fused optimizer on its shard (csrc/shard.cu).  Two stream-ordered collectives per step act as the only barriers.

Usage (every rank):
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", LOCAL_RANK))
    model = xDeepFM(cols, cols, device="cuda:%d" % LOCAL_RANK)
    model.distribute()              # shards the tables, broadcasts the dense parameters of rank 0
    model.compile("adam", "binary_crossentropy")
    model.fit(x, y, batch_size=per_gpu_batch)   # every rank passes the same arrays; rank r trains on its slice of each batch

The integer routing functions at the top are pure Python so that the CPU test-suite (gloo, world_size 2) covers them.
"""
import ctypes
import os

import torch

from . import _native as N
from . import ops


# ------------------------------------------------------------------------------------------------
# routing: pure integer logic
# ------------------------------------------------------------------------------------------------
def owner_of(row, G):
    """Rank that stores global row `row`."""
    return row % G


def local_row(row, G):
    """Row index inside the owner's shard of the table."""
    return row // G


def shard_rows(V, rank, G):
    """Number of rows of a V-row table stored on `rank` (rows r with r % G == rank)."""
    return max(0, (V - rank + G - 1) // G)


def shard_layout(rows, G):
    """rows[t] = rows of table t.  Returns (base, total): base[g][t] = first local row of table t inside rank g's contiguous
    shard buffer, total[g] = rows of that buffer."""
    base, total = [], []
    for g in range(G):
        off, b = 0, []
        for V in rows:
            b.append(off)
            off += shard_rows(V, g, G)
        base.append(b)
        total.append(off)
    return base, total


def key_stride(total, G):
    """Owner stride S of the backward key space: key = owner * S + local row.  Power of two > every rank's row count
    (S itself is the sentinel of the owner-side merge sort)."""
    S = 1
    while S <= max(total):
        S <<= 1
    if S * G > (1 << 32):
        raise ValueError("row-sharded tables: %d ranks x %d rows per rank exceed the 32-bit key space" % (G, max(total)))
    return S


def shard_key(table, row, base, G, S):
    """Backward sort key of (table, global row)."""
    g = owner_of(row, G)
    return g * S + base[g][table] + local_row(row, G)


def take_shard(full, rank, G):
    """Rows of a full [V, w] table owned by `rank`."""
    return full[rank::G]


def merge_shards(shards, V):
    """Inverse of take_shard: list of G shards -> full [V, w] table."""
    G = len(shards)
    full = shards[0].new_empty((V,) + tuple(shards[0].shape[1:]))
    for g, s in enumerate(shards):
        full[g::G] = s[:shard_rows(V, g, G)]
    return full


def rank_slice(lo, hi, rank, G):
    """Contiguous sub-range of the global batch [lo, hi) trained by `rank` (sizes differ by at most one)."""
    n = hi - lo
    q, rem = divmod(n, G)
    a = lo + rank * q + min(rank, rem)
    return a, a + q + (1 if rank < rem else 0)


# ------------------------------------------------------------------------------------------------
# peer-mappable device memory
# ------------------------------------------------------------------------------------------------
class _RawCuda:
    """Exposes a raw device allocation to torch through __cuda_array_interface__ (torch keeps this object alive)."""

    def __init__(self, owner, ptr, nbytes):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerBuffer:
    """A cudaMalloc'ed, zero-filled buffer on the current device that the other ranks of the node can map (cudaIpc)."""

    def __init__(self, nbytes, device):
        self.device = torch.device(device)
        self.nbytes = int(max(nbytes, 256))
        p = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            N.check(N.lib().xdfm_ipc_alloc(self.nbytes, ctypes.byref(p)))
        self.ptr = p.value
        self._opened = []
        self.bytes_view = torch.as_tensor(_RawCuda(self, self.ptr, self.nbytes), device=self.device)

    def handle(self):
        h = (ctypes.c_char * 64)()
        N.check(N.lib().xdfm_ipc_export(ctypes.c_void_p(self.ptr), h))
        return bytes(h)

    def view(self, offset, shape, dtype):
        n = 1
        for s in shape:
            n *= int(s)
        nb = n * torch.empty((), dtype=dtype).element_size()
        return self.bytes_view[offset:offset + nb].view(dtype).view(*shape)

    def open_peer(self, handle):
        p = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            N.check(N.lib().xdfm_ipc_open(ctypes.c_char_p(handle), ctypes.byref(p)))
        self._opened.append(p.value)
        return p.value

    def close(self):
        for p in self._opened:
            N.lib().xdfm_ipc_close(ctypes.c_void_p(p))
        self._opened = []

    def __del__(self):
        try:
            self.close()
            if self.ptr:
                N.lib().xdfm_ipc_free(ctypes.c_void_p(self.ptr))
                self.ptr = 0
        except Exception:
            pass


def _align(x, a=256):
    return (x + a - 1) // a * a


def _ptr_array(vals):
    arr = (ctypes.c_void_p * len(vals))()
    for i, v in enumerate(vals):
        arr[i] = v
    return arr


# ------------------------------------------------------------------------------------------------
# the sharded tables of one rank
# ------------------------------------------------------------------------------------------------
class ShardedSparse:
    """Embedding shard [rows_g, D] + first-order shard [rows_g] of one rank, the exchange buffers of the backward, and the
    peer pointers.  `connect()` wires G instances together -- across processes with cudaIpc handles, or inside one process
    (tests: G emulated ranks on one GPU, kernels run one after another) with raw pointers."""

    def __init__(self, rank, G, table_of, rows, vocab, D, device):
        self.rank, self.G, self.D = int(rank), int(G), int(D)
        self.table_of, self.rows, self.vocab = list(table_of), [int(r) for r in rows], [int(v) for v in vocab]
        self.m = len(self.table_of)
        if self.m < 1 or self.m > N.MAX_FIELDS:
            raise ValueError("row-sharded tables need 1..%d sparse features" % N.MAX_FIELDS)
        if D % 4 != 0:
            raise ValueError("row-sharded tables need embedding_dim %% 4 == 0 (128-bit rows), got %d" % D)
        self.device = torch.device(device)
        self.base, self.total = shard_layout(self.rows, self.G)
        self.S = key_stride(self.total, self.G)
        self.local_rows = self.total[self.rank]
        # one exported allocation: tables [local_rows, D] / [local_rows], both optimizer moments of each, and the per-row
        # "current up to step" counters of the lazy table semantics (peers read all of them when they replay a stale row)
        nr = max(self.local_rows, 1)
        sizes = [("emb", nr * self.D * 4), ("lin", nr * 4), ("s1", nr * self.D * 4), ("s1_lin", nr * 4), ("s2", nr * self.D * 4),
                 ("s2_lin", nr * 4), ("last", nr * 4), ("last_lin", nr * 4)]
        self._off, off = {}, 0
        for name, nb in sizes:
            self._off[name] = off
            off += _align(nb)
        self._emb_off, self._lin_off = self._off["emb"], self._off["lin"]
        self.tables = PeerBuffer(off, self.device)
        self.emb = self.tables.view(self._off["emb"], (self.local_rows, self.D), torch.float32)
        self.lin = self.tables.view(self._off["lin"], (self.local_rows, 1), torch.float32)
        self._s1_buf = self.tables.view(self._off["s1"], (self.local_rows, self.D), torch.float32)
        self._s1_lin_buf = self.tables.view(self._off["s1_lin"], (self.local_rows, 1), torch.float32)
        self._s2_buf = self.tables.view(self._off["s2"], (self.local_rows, self.D), torch.float32)
        self._s2_lin_buf = self.tables.view(self._off["s2_lin"], (self.local_rows, 1), torch.float32)
        self.last = self.tables.view(self._off["last"], (max(self.local_rows, 1),), torch.int32)
        self.last_lin = self.tables.view(self._off["last_lin"], (max(self.local_rows, 1),), torch.int32)
        self.n_states = 2           # set by the optimizer (0 sgd, 1 adagrad/rmsprop, 2 adam)
        self.lazy = None            # set by the optimizer: dict(cfg_emb, cfg_lin, opt_dev, hist, hist_base) -> lookups replay stale rows
        self.peer_ptrs = None       # device int64 [8 * G]
        self.feat_base = torch.tensor([[self.base[g][t] for t in self.table_of] for g in range(self.G)], dtype=torch.int64,
                                      device=self.device).reshape(-1).contiguous()
        self._c_vocab = N.i32_array(self.vocab)
        self.cap = 0                    # keys per rank the exchange buffers can hold
        self.exchange = None
        self.peer_emb = self.peer_lin = None          # device arrays [G] of pointers
        self._peer_tables = None
        self._peer_x = None
        self.stash = {}
        # XDFM_SHARD_UNIQUE=0: every lookup reads its row from the owner (the first design; kept for batches beyond the exchange capacity)
        self.unique_lookup = os.environ.get("XDFM_SHARD_UNIQUE", "1") != "0"
        self._seg_key = self._uniq = None
        self.anchor = torch.zeros(1, device=self.device, requires_grad=True)
        # optimizer state (allocated by FusedOptimizer.prepare)
        self.s1 = self.s2 = self.s1_lin = self.s2_lin = None
        self.bitmap = torch.zeros((self.local_rows + 31) // 32 + 1, dtype=torch.int32, device=self.device)
        self._row_off = N.i64_array([0, self.local_rows])

    # ---- wiring -----------------------------------------------------------------------------------
    def _exchange_layout(self, cap):
        o = {}
        off = 0
        o["keys"] = off; off += _align(cap * 4)
        o["gsum"] = off; off += _align(cap * self.D * 4)
        o["gsum_lin"] = off; off += _align(cap * 4)
        o["ranges"] = off; off += _align((self.G + 1) * 4)
        o["total"] = off
        return o

    def alloc_exchange(self, cap):
        """(Re)allocate the exported exchange buffers for `cap` keys per rank.  Collective: follow with connect()."""
        self.cap = int(max(cap, 1))
        self._xl = self._exchange_layout(self.cap)
        self.exchange = PeerBuffer(self._xl["total"], self.device)
        x = self.exchange
        self.x_keys = x.view(self._xl["keys"], (self.cap,), torch.int32)
        self.x_gsum = x.view(self._xl["gsum"], (self.cap, self.D), torch.float32)
        self.x_gsum_lin = x.view(self._xl["gsum_lin"], (self.cap,), torch.float32)
        self.x_ranges = x.view(self._xl["ranges"], (self.G + 1,), torch.int32)
        n_cap = self.cap * self.G
        dev = self.device
        self.seg_off = torch.empty(self.cap + 1, dtype=torch.int32, device=dev)
        self.pos = torch.empty(self.cap, dtype=torch.int32, device=dev)
        self.nseg = torch.zeros(1, dtype=torch.int32, device=dev)
        self.p_rows = torch.empty((n_cap, self.D), dtype=torch.float32, device=dev)
        self.p_rows_lin = torch.empty(n_cap, dtype=torch.float32, device=dev)
        self.p_uniq = torch.empty(n_cap, dtype=torch.int32, device=dev)
        self.p_seg_off = torch.empty(n_cap + 1, dtype=torch.int32, device=dev)
        self.p_pos = torch.empty(n_cap, dtype=torch.int32, device=dev)
        self.p_nseg = torch.zeros(1, dtype=torch.int32, device=dev)
        self.p_gsum = torch.empty((n_cap, self.D), dtype=torch.float32, device=dev)
        self.p_gsum_lin = torch.empty(n_cap, dtype=torch.float32, device=dev)
        self.ws = torch.empty(int(N.lib().xdfm_shard_workspace_bytes(n_cap)), dtype=torch.uint8, device=dev)
        # lookup through the batch's distinct rows: compact rows fetched over NVLink + the segment of every lookup
        self.u_emb = torch.empty((self.cap, self.D), dtype=torch.float32, device=dev)
        self.u_lin = torch.empty(self.cap, dtype=torch.float32, device=dev)
        self.inv = torch.empty(self.cap, dtype=torch.int32, device=dev)
        self._seg_key = self._uniq = None

    def export(self):
        """Picklable description of this rank's exported buffers (all_gather_object it, then connect())."""
        return {"rank": self.rank, "tables": self.tables.handle(), "lin_off": self._lin_off, "off": dict(self._off),
                "exchange": self.exchange.handle() if self.exchange is not None else None, "xl": getattr(self, "_xl", None)}

    def local_pointers(self):
        """Same description with raw pointers (in-process wiring)."""
        return {"rank": self.rank, "tables_ptr": self.tables.ptr, "lin_off": self._lin_off, "off": dict(self._off),
                "exchange_ptr": self.exchange.ptr if self.exchange is not None else None, "xl": getattr(self, "_xl", None)}

    def connect(self, infos):
        """infos[g] = export() or local_pointers() of rank g."""
        assert len(infos) == self.G
        tp, xp = [], []
        for g, info in enumerate(infos):
            if g == self.rank:
                tp.append(self.tables.ptr)
                xp.append(self.exchange.ptr if self.exchange is not None else 0)
            elif "tables_ptr" in info:
                tp.append(info["tables_ptr"])
                xp.append(info["exchange_ptr"] or 0)
            else:
                if self._peer_tables is not None and self._peer_tables[g]:
                    tp.append(self._peer_tables[g])          # tables stay mapped across exchange re-allocations
                else:
                    tp.append(self.tables.open_peer(info["tables"]))
                xp.append(self.exchange.open_peer(info["exchange"]) if info["exchange"] is not None else 0)
        self._peer_tables, self._peer_x = tp, xp
        dev = self.device
        self.peer_emb = torch.tensor(tp, dtype=torch.int64, device=dev)
        self.peer_lin = torch.tensor([tp[g] + infos[g]["lin_off"] for g in range(self.G)], dtype=torch.int64, device=dev)
        self._peer_tp, self._peer_off = tp, [info["off"] for info in infos]
        self._build_peer_ptrs()
        if self.exchange is not None:
            xls = [info["xl"] for info in infos]
            self._c_peer_keys = _ptr_array([xp[g] + xls[g]["keys"] for g in range(self.G)])
            self._c_peer_gsum = _ptr_array([xp[g] + xls[g]["gsum"] for g in range(self.G)])
            self._c_peer_gsum_lin = _ptr_array([xp[g] + xls[g]["gsum_lin"] for g in range(self.G)])
            self._c_peer_ranges = _ptr_array([xp[g] + xls[g]["ranges"] for g in range(self.G)])

    def _build_peer_ptrs(self):
        """[8 * G] addresses: emb, lin, s1, s1_lin, s2, s2_lin, last, last_lin of every rank's shard (missing moments are 0)."""
        names = ["emb", "lin", "s1", "s1_lin", "s2", "s2_lin", "last", "last_lin"]
        need = {"emb": 0, "lin": 0, "s1": 1, "s1_lin": 1, "s2": 2, "s2_lin": 2, "last": 0, "last_lin": 0}
        vals = []
        for g in range(self.G):
            for nme in names:
                vals.append(self._peer_tp[g] + self._peer_off[g][nme] if self.n_states >= need[nme] else 0)
        self.peer_ptrs = torch.tensor(vals, dtype=torch.int64, device=self.device)

    def set_states(self, n_states):
        """Called by the optimizer: which moment buffers exist (0 sgd, 1 adagrad / rmsprop, 2 adam)."""
        self.n_states = int(n_states)
        self.s1, self.s1_lin = (self._s1_buf, self._s1_lin_buf) if n_states >= 1 else (None, None)
        self.s2, self.s2_lin = (self._s2_buf, self._s2_lin_buf) if n_states >= 2 else (None, None)
        if getattr(self, "_peer_tp", None) is not None:
            self._build_peer_ptrs()

    def optimizer_state(self):
        """This rank's moments (call after the optimizer flushed: every row is current)."""
        return {"rank": self.rank, "G": self.G, "local_rows": self.local_rows,
                "s1": None if self.s1 is None else self.s1.detach().clone(), "s1_lin": None if self.s1_lin is None else self.s1_lin.detach().clone(),
                "s2": None if self.s2 is None else self.s2.detach().clone(), "s2_lin": None if self.s2_lin is None else self.s2_lin.detach().clone()}

    def load_optimizer_state(self, st, steps):
        if (st["rank"], st["G"], st["local_rows"]) != (self.rank, self.G, self.local_rows):
            raise ValueError("shard optimizer state of rank %d/%d (%d rows) does not fit rank %d/%d (%d rows)" % (
                st["rank"], st["G"], st["local_rows"], self.rank, self.G, self.local_rows))
        for mine, key in ((self.s1, "s1"), (self.s1_lin, "s1_lin"), (self.s2, "s2"), (self.s2_lin, "s2_lin")):
            if mine is not None and st[key] is not None:
                mine.copy_(st[key])
        self.last.fill_(int(steps))
        self.last_lin.fill_(int(steps))

    # ---- forward ------------------------------------------------------------------------------------
    def _segments_of(self, ids):
        """Sort this batch's keys owner-major and run-length encode them: (x_keys, x_ranges) in the exchange buffer, (seg_off, pos,
        nseg) next to it.  Run once per batch, before the lookup; the backward (reduce_local) reuses the result."""
        B, m = ids.shape
        with ops.timed("embed_segments"):
            N.check(N.lib().xdfm_shard_segments(N.ptr(ids), B, m, self.G, self.S, N.ptr(self.feat_base), self._c_vocab, N.ptr(self.ws),
                                                self.ws.numel(), N.ptr(self.x_keys), N.ptr(self.seg_off), N.ptr(self.pos), N.ptr(self.nseg),
                                                N.ptr(self.x_ranges), N.stream_ptr()))
        self._seg_key = (ids, ids._version)           # strong reference: the address cannot be recycled while cached
        self._uniq = None

    def _segments_current(self, ids):
        k = self._seg_key
        return k is not None and k[0] is ids and k[1] == ids._version

    def tables_changed(self):
        """Rows were rewritten (optimizer step, catch-up, flush, load): fetched copies of them are stale."""
        self._uniq = None

    def _gather_unique(self, ids, out, lin, dense, dense_w, nd, lin_rows=None):
        """Every DISTINCT row of the batch crosses NVLink once (replayed if stale) into (u_emb, u_lin); the per-sample tensors are
        expanded from them locally.  One fetch serves the embedding lookup and the first-order term of the same forward."""
        B, m = ids.shape
        L = N.lib()
        st = N.stream_ptr()
        need = ([] if out is None else ["emb"]) + ([] if lin is None and lin_rows is None else ["lin"])
        if self._uniq is None or not self._segments_current(ids) or any(p in self._uniq for p in need):
            self._segments_of(ids)
            lz = self.lazy
            def fetch(u_emb, u_lin, inv):
                N.check(L.xdfm_embed_fetch_unique_sharded(
                    N.ptr(self.peer_ptrs), self.G, self.S, self.D, N.ptr(self.x_keys), N.ptr(self.seg_off), N.ptr(self.pos), N.ptr(self.nseg),
                    B * m, lz["cfg_emb"] if lz else None, lz["cfg_lin"] if lz else None, N.ptr(lz["opt_dev"]) if lz else None,
                    N.ptr(lz["hist"]) if lz else None, lz["hist_base"] if lz else 0, N.ptr(u_emb), N.ptr(u_lin), N.ptr(inv), st))

            if ops.TIMERS is not None:          # operator-timing pass: the three kernels one by one
                with ops.timed("embed_gather"):
                    fetch(self.u_emb, None, None)
                with ops.timed("embed_gather_lin"):
                    fetch(None, self.u_lin, None)
                with ops.timed("embed_expand"):
                    fetch(None, None, self.inv)
            else:
                fetch(self.u_emb, self.u_lin, self.inv)
            self._uniq = set()
        with ops.timed("embed_expand"):
            if out is not None or lin is not None:
                N.check(L.xdfm_embed_expand_unique(N.ptr(self.u_emb), N.ptr(self.u_lin), N.ptr(self.inv), B, m, self.D, N.ptr(out),
                                                   N.ptr(dense) if nd > 0 else None, nd, N.ptr(dense_w) if nd > 0 else None, N.ptr(lin), st))
            if lin_rows is not None:
                N.check(L.xdfm_embed_expand_unique_lin_rows(N.ptr(self.u_lin), N.ptr(self.inv), B * m, N.ptr(lin_rows), st))
        self._uniq.update(need)
        return out, lin

    def gather_lin_rows(self, ids):
        """[B, m, 1] first-order row of every lookup, un-summed (models with multi-value features pool them per field).  Always
        through the distinct rows of the batch; batches beyond the exchange capacity are looked up in slices."""
        B, m = ids.shape
        assert m == self.m
        rows = torch.empty((B, m, 1), dtype=torch.float32, device=ids.device)
        if B == 0:
            return rows
        if self.exchange is None or self.peer_ptrs is None:
            raise RuntimeError("row-sharded tables are not connected yet (distribute() wires them)")
        step = max(self.cap // m, 1)
        if step * m > self.cap:
            raise RuntimeError("row-sharded tables: exchange capacity %d is below one sample's %d lookups" % (self.cap, m))
        if B <= step:
            self._gather_unique(ids, None, None, None, None, 0, lin_rows=rows)
        else:
            for a in range(0, B, step):
                self._gather_unique(ids[a:a + step], None, None, None, None, 0, lin_rows=rows[a:a + step])
        return rows

    def gather(self, ids, want_emb=True, dense=None, dense_w=None, want_lin=False):
        B, m = ids.shape
        assert m == self.m
        dev = ids.device
        out = torch.empty((B, m, self.D), dtype=torch.float32, device=dev) if want_emb else None
        lin = torch.empty((B,), dtype=torch.float32, device=dev) if want_lin else None
        nd = 0 if dense is None or dense_w is None else dense.shape[1]
        if self.unique_lookup and self.exchange is not None and 0 < B * m <= self.cap and self.peer_ptrs is not None:
            return self._gather_unique(ids, out, lin, dense, dense_w, nd)
        if self.lazy is not None:
            lz = self.lazy
            with ops.timed("embed_gather"):
                N.check(N.lib().xdfm_embed_gather_sharded_lazy(N.ptr(self.peer_ptrs), N.ptr(self.feat_base), self._c_vocab, N.ptr(ids), B, m,
                                                               self.D, self.G, lz["cfg_emb"], lz["cfg_lin"], N.ptr(lz["opt_dev"]),
                                                               N.ptr(lz["hist"]), lz["hist_base"], N.ptr(out),
                                                               N.ptr(dense) if nd > 0 else None, nd, N.ptr(dense_w) if nd > 0 else None,
                                                               N.ptr(lin), N.stream_ptr()))
            return out, lin
        with ops.timed("embed_gather"):
            N.check(N.lib().xdfm_embed_gather_sharded(N.ptr(self.peer_emb), N.ptr(self.peer_lin), N.ptr(self.feat_base), self._c_vocab,
                                                      N.ptr(ids), B, m, self.D, self.G, N.ptr(out), N.ptr(dense) if nd > 0 else None,
                                                      nd, N.ptr(dense_w) if nd > 0 else None, N.ptr(lin), N.stream_ptr()))
        return out, lin

    # ---- backward, batch side -----------------------------------------------------------------------
    def reduce_local(self):
        """Sort this rank's batch keys owner-major, reduce duplicate rows and leave (keys, sums, owner ranges) in the exchange
        buffers.  Uses the tensors stashed by the backward of ShardedGather / ShardedLinearTerm."""
        ids = self.stash.get("ids")
        st = N.stream_ptr()
        L = N.lib()
        if ids is None or ids.shape[0] == 0:
            self.nseg.zero_()
            self.x_ranges.zero_()
            self.stash = {}
            return
        B, m = ids.shape
        n = B * m
        if n > self.cap:
            raise RuntimeError("row-sharded tables: batch of %d keys exceeds the exchange capacity %d; call "
                               "model.distribute(max_batch=...) with the largest per-GPU batch" % (n, self.cap))
        if not self._segments_current(ids):             # (the unique-row lookup of the forward already sorted this batch)
            self._segments_of(ids)
        self._seg_key = None                             # one forward, one backward
        with ops.timed("embed_scatter"):
            demb, dlin, dlin_rows = self.stash.get("demb"), self.stash.get("dlin"), self.stash.get("dlin_rows")
            if demb is None:
                self.x_gsum[:n].zero_()
            if dlin is None and dlin_rows is None:
                self.x_gsum_lin[:n].zero_()
            if demb is not None or dlin is not None:
                N.check(L.xdfm_embed_bwd_reduce(N.ptr(demb), N.ptr(dlin), N.ptr(self.pos), N.ptr(self.seg_off), N.ptr(self.nseg), n, m,
                                                self.D, N.ptr(self.x_gsum) if demb is not None else None,
                                                N.ptr(self.x_gsum_lin) if dlin is not None else None, st))
            if dlin_rows is not None:
                # one gradient per LOOKUP (multi-value features: pooled first-order rows), reduced like a width-1 embedding
                N.check(L.xdfm_embed_bwd_reduce(N.ptr(dlin_rows), None, N.ptr(self.pos), N.ptr(self.seg_off), N.ptr(self.nseg), n, m, 1,
                                                N.ptr(self.x_gsum_lin), None, st))
        self.stash = {}

    # ---- backward, owner side -----------------------------------------------------------------------
    def pull_segments(self):
        """Pull this rank's ranges from all peers and merge them: afterwards (p_uniq, p_gsum, p_gsum_lin, p_nseg) hold the
        unique local rows touched by ANY rank's batch and their summed gradients."""
        L = N.lib()
        st = N.stream_ptr()
        n_cap = self.cap * self.G
        with ops.timed("embed_pull"):
            N.check(L.xdfm_shard_pull_segments(self._c_peer_keys, self._c_peer_gsum, self._c_peer_gsum_lin, self._c_peer_ranges, self.G,
                                               self.rank, self.S, self.D, n_cap, N.ptr(self.ws), self.ws.numel(), N.ptr(self.p_rows),
                                               N.ptr(self.p_rows_lin), N.ptr(self.p_uniq), N.ptr(self.p_seg_off), N.ptr(self.p_pos),
                                               N.ptr(self.p_nseg), st))
            N.check(L.xdfm_embed_bwd_reduce(N.ptr(self.p_rows), N.ptr(self.p_rows_lin), N.ptr(self.p_pos), N.ptr(self.p_seg_off),
                                            N.ptr(self.p_nseg), n_cap, 1, self.D, N.ptr(self.p_gsum), N.ptr(self.p_gsum_lin), st))

    def catch_up_pulled(self, cfg_emb, cfg_lin, opt_dev, hist, hist_base, reg_out):
        """Lazy semantics, owner side: the rows about to be updated (p_uniq) are replayed up to the completed-step count first.
        Must run BEFORE the step counter is advanced."""
        self.tables_changed()
        L = N.lib()
        st = N.stream_ptr()
        n_cap = self.cap * self.G
        with ops.timed("rows_opt"):
            for cfg, w, s1, s2, width, last in ((cfg_emb, self.emb, self.s1, self.s2, self.D, self.last),
                                                (cfg_lin, self.lin, self.s1_lin, self.s2_lin, 1, self.last_lin)):
                N.check(L.xdfm_rows_catchup(cfg, N.ptr(opt_dev), N.ptr(hist), hist_base, N.ptr_array([w]),
                                            N.ptr_array([s1]) if s1 is not None else None, N.ptr_array([s2]) if s2 is not None else None,
                                            N.ptr(last), self._row_off, 1, width, N.ptr(self.p_uniq), N.ptr(self.p_nseg), n_cap,
                                            N.ptr(reg_out), st))

    def mark_pulled(self, opt_dev):
        for last in (self.last, self.last_lin):
            N.check(N.lib().xdfm_rows_mark_current(N.ptr(last), N.ptr(self.p_uniq), N.ptr(self.p_nseg), self.cap * self.G, N.ptr(opt_dev),
                                                   N.stream_ptr()))

    def flush_rows(self, cfg_emb, cfg_lin, opt_dev, hist, hist_base, reg_out):
        self.tables_changed()
        L = N.lib()
        st = N.stream_ptr()
        if self.local_rows == 0:
            return
        with ops.timed("rows_opt"):
            for cfg, w, s1, s2, width, last in ((cfg_emb, self.emb, self.s1, self.s2, self.D, self.last),
                                                (cfg_lin, self.lin, self.s1_lin, self.s2_lin, 1, self.last_lin)):
                N.check(L.xdfm_rows_flush(cfg, N.ptr(opt_dev), N.ptr(hist), hist_base, N.ptr_array([w]),
                                          N.ptr_array([s1]) if s1 is not None else None, N.ptr_array([s2]) if s2 is not None else None,
                                          N.ptr(last), self._row_off, 1, width, N.ptr(reg_out), st))

    def apply_optimizer(self, cfg_emb, cfg_lin, opt_dev, grad_scale, dense_pass, reg_out):
        self.tables_changed()
        L = N.lib()
        st = N.stream_ptr()
        n_cap = self.cap * self.G
        with ops.timed("rows_opt"):
            for cfg, w, s1, s2, width, gs in ((cfg_emb, self.emb, self.s1, self.s2, self.D, self.p_gsum),
                                              (cfg_lin, self.lin, self.s1_lin, self.s2_lin, 1, self.p_gsum_lin)):
                dp = dense_pass
                if cfg.kind == N.OPT["sgd"] and cfg.l2 == 0.0:
                    dp = 0
                N.check(L.xdfm_rows_opt(cfg, N.ptr(opt_dev), N.ptr_array([w]), N.ptr_array([s1]) if s1 is not None else None,
                                        N.ptr_array([s2]) if s2 is not None else None, self._row_off, 1, width, N.ptr(self.p_uniq),
                                        N.ptr(gs), N.ptr(self.p_nseg), n_cap, float(grad_scale), N.ptr(self.bitmap), dp,
                                        N.ptr(reg_out), st))


class ShardedGather(torch.autograd.Function):
    """[B, m, D] lookup in the row-sharded tables; the backward only stashes d(out) -- ShardedSparse.reduce_local() (called by the
    fused optimizer) turns it into per-row sums."""

    @staticmethod
    def forward(ctx, sh, ids, anchor):
        ops.require_cuda(ids, "ShardedGather")
        out, _ = sh.gather(ids, want_emb=True)
        ctx.sh, ctx.ids = sh, ids
        return out

    @staticmethod
    def backward(ctx, dout):
        ctx.sh.stash["ids"] = ctx.ids
        ctx.sh.stash["demb"] = ops._f32c(dout)
        return None, None, None


class ShardedLinearTerm(torch.autograd.Function):
    """First-order term over the row-sharded [V,1] tables + dense @ weight -> [B, 1]."""

    @staticmethod
    def forward(ctx, sh, ids, dense, dense_w, anchor):
        ops.require_cuda(ids, "ShardedLinearTerm")
        _, lin = sh.gather(ids, want_emb=False, dense=dense, dense_w=dense_w, want_lin=True)
        ctx.sh, ctx.ids, ctx.dense = sh, ids, dense
        ctx.nd = 0 if dense is None or dense_w is None else dense.shape[1]
        ctx.dense_w_shape = None if dense_w is None else dense_w.shape
        return lin.view(-1, 1)

    @staticmethod
    def backward(ctx, dout):
        dlin = ops._f32c(dout).reshape(-1)
        B = dlin.shape[0]
        d_dense_w = None
        if ctx.nd > 0:
            d_dense_w = torch.empty(ctx.nd, dtype=torch.float32, device=dlin.device)
            ws = ops.workspace("wcolsum", N.lib().xdfm_wcolsum_workspace_bytes(ctx.nd), dlin.device)
            N.check(N.lib().xdfm_wcolsum(N.ptr(ctx.dense), B, ctx.nd, ctx.nd, N.ptr(dlin), N.ptr(d_dense_w), 0, N.ptr(ws), ws.numel(),
                                         N.stream_ptr()))
            d_dense_w = d_dense_w.view(ctx.dense_w_shape)
        ctx.sh.stash["ids"] = ctx.ids
        ctx.sh.stash["dlin"] = dlin
        return None, None, None, d_dense_w, None


class ShardedLinearRows(torch.autograd.Function):
    """[B, m, 1] first-order rows of the row-sharded [V,1] tables, one per lookup (the model pools the positions of multi-value
    features per field before summing: basemodel.py:63-92 with inputs.py:141-155)."""

    @staticmethod
    def forward(ctx, sh, ids, anchor):
        ops.require_cuda(ids, "ShardedLinearRows")
        ctx.sh, ctx.ids = sh, ids
        return sh.gather_lin_rows(ids)

    @staticmethod
    def backward(ctx, dout):
        ctx.sh.stash["ids"] = ctx.ids
        ctx.sh.stash["dlin_rows"] = ops._f32c(dout).reshape(-1)
        return None, None, None


# ------------------------------------------------------------------------------------------------
# process-group context of a distributed model
# ------------------------------------------------------------------------------------------------
class DistContext:
    def __init__(self, group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("model.distribute(): call torch.distributed.init_process_group('nccl') first (one process per GPU)")
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.sharded = None
        self._token = None

    def all_gather_object(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def broadcast(self, t, src=0):
        if self.world > 1:
            self.dist.broadcast(t, src=self.dist.get_global_rank(self.group, src) if self.group is not None else src, group=self.group)
        return t

    def barrier(self, device):
        """Stream-ordered barrier: a one-element all-reduce on the compute stream."""
        if self.world > 1:
            if self._token is None or self._token.device != torch.device(device):
                self._token = torch.zeros(1, dtype=torch.float32, device=device)
            self.dist.all_reduce(self._token, group=self.group)

    def ensure_capacity(self, n_keys):
        """Collective: (re)allocate + re-wire the exchange buffers so that every rank can post `n_keys` keys per step."""
        sh = self.sharded
        if sh.exchange is not None and n_keys <= sh.cap:
            return
        torch.cuda.synchronize(sh.device)
        sh.alloc_exchange(n_keys)
        sh.connect(self.all_gather_object(sh.export()))
        torch.cuda.synchronize(sh.device)
        self.barrier(sh.device)
        torch.cuda.synchronize(sh.device)


def attach(model, group=None, max_batch=None):
    """Turn `model` (a BaseModel on a CUDA device) into this rank's part of a hybrid-parallel model.  Collective."""
    dev = torch.device(model.device)
    if dev.type != "cuda":
        raise RuntimeError("distribute(): the xdeepfm-b200 path needs device='cuda:N'; no CPU fallback")
    ctx = DistContext(group)
    lin_names = [fc.name for fc in list(model.linear_model.sparse_feature_columns) + list(model.linear_model.varlen_sparse_feature_columns)]
    dnn_names = [fc.name for fc in list(model._dnn_sparse) + list(model._dnn_varlen)]
    if lin_names != dnn_names or model._dnn_sparse_sel != model._lin_sparse_sel or not dnn_names:
        raise NotImplementedError("distribute(): the linear and the deep part must use the same sparse feature columns "
                                  "(as every xdftrain*.py script builds them)")
    emb_plan, lin_plan = model._emb_plan, model.linear_model._plan
    if emb_plan.table_of != lin_plan.table_of or emb_plan.rows != lin_plan.rows:
        raise NotImplementedError("distribute(): embedding and first-order tables must share one vocabulary layout")
    emb_tables = [e.weight for e in model.embedding_dict.values()]
    lin_tables = [e.weight for e in model.linear_model.embedding_dict.values()]
    table_ids = set(id(p) for p in emb_tables + lin_tables)
    with torch.no_grad():
        for p in model.parameters():
            if id(p) not in table_ids:
                ctx.broadcast(p.data, 0)
        sh = ShardedSparse(ctx.rank, ctx.world, emb_plan.table_of, emb_plan.rows, emb_plan.vocab, emb_plan.width, dev)
        emb_mods = list(model.embedding_dict.values())
        ctx.deferred = any(hasattr(e, "deferred_rows") for e in emb_mods)
        gen = torch.Generator(device=dev)
        for t, (we, wl) in enumerate(zip(emb_tables, lin_tables)):
            a = sh.base[ctx.rank][t]
            k = shard_rows(sh.rows[t], ctx.rank, ctx.world)
            if hasattr(emb_mods[t], "deferred_rows"):
                # built under deferred_tables(): the table never existed as a whole; this rank's rows are initialised in place
                # (N(0, init_std) as inputs.create_embedding_matrix does), seeded per (table, rank)
                std = float(getattr(emb_mods[t], "init_std", 1e-4))
                gen.manual_seed(1024 + 7919 * t + ctx.rank)
                sh.emb[a:a + k].normal_(0.0, std, generator=gen)
                sh.lin[a:a + k].normal_(0.0, std, generator=gen)
            else:
                ctx.broadcast(we.data, 0)
                ctx.broadcast(wl.data, 0)
                sh.emb[a:a + k].copy_(take_shard(we.data, ctx.rank, ctx.world))
                sh.lin[a:a + k].copy_(take_shard(wl.data, ctx.rank, ctx.world))
            # the full tables are not kept: state_dict() re-assembles them from the shards
            we.data = torch.empty((0, we.shape[1]), dtype=we.dtype, device=dev)
            wl.data = torch.empty((0, 1), dtype=wl.dtype, device=dev)
    ctx.sharded = sh
    model._dist = ctx
    sh.alloc_exchange((max_batch or 256) * sh.m)
    sh.connect(ctx.all_gather_object(sh.export()))
    torch.cuda.synchronize(dev)
    ctx.barrier(dev)
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()
    return ctx


def gather_tables(model):
    """Collective: {state_dict key: full table} re-assembled from all shards (reference key layout)."""
    ctx = model._dist
    sh = ctx.sharded
    out = {}
    total_bytes = sum(sh.rows) * (sh.D + 1) * 4
    if getattr(ctx, "deferred", False) and total_bytes > (8 << 30):
        raise RuntimeError("state_dict(): the tables of this model (%.1f GB) were built sharded (deferred_tables) and are not re-assembled "
                           "on one device; use deepctr.distributed.sharded_state_dict(model) (per-rank shards)" % (total_bytes / 1e9))
    emb_names = list(model.embedding_dict.keys())
    lin_names = list(model.linear_model.embedding_dict.keys())
    for t, V in enumerate(sh.rows):
        for buf, key, width in ((sh.emb, "embedding_dict.%s.weight" % emb_names[t], sh.D),
                                (sh.lin, "linear_model.embedding_dict.%s.weight" % lin_names[t], 1)):
            kmax = shard_rows(V, 0, ctx.world)
            mine = torch.zeros((kmax, width), dtype=torch.float32, device=sh.device)
            a = sh.base[ctx.rank][t]
            k = shard_rows(V, ctx.rank, ctx.world)
            mine[:k].copy_(buf[a:a + k])
            parts = [torch.empty_like(mine) for _ in range(ctx.world)]
            if ctx.world > 1:
                ctx.dist.all_gather(parts, mine, group=ctx.group)
            else:
                parts = [mine]
            out[key] = merge_shards(parts, V)
    return out


def scatter_tables(model, state):
    """Inverse of gather_tables for load_state_dict: every rank copies its rows out of the full tables in `state`."""
    ctx = model._dist
    sh = ctx.sharded
    emb_names = list(model.embedding_dict.keys())
    lin_names = list(model.linear_model.embedding_dict.keys())
    used = []
    with torch.no_grad():
        for t, V in enumerate(sh.rows):
            for buf, key in ((sh.emb, "embedding_dict.%s.weight" % emb_names[t]),
                             (sh.lin, "linear_model.embedding_dict.%s.weight" % lin_names[t])):
                if key not in state:
                    continue
                full = state[key]
                if full.shape[0] != V:
                    raise RuntimeError("size mismatch for %s: checkpoint has %d rows, model has %d" % (key, full.shape[0], V))
                a = sh.base[ctx.rank][t]
                k = shard_rows(V, ctx.rank, ctx.world)
                buf[a:a + k].copy_(take_shard(full, ctx.rank, ctx.world).to(sh.device, torch.float32).reshape(k, buf.shape[1]))
                used.append(key)
    return used


def sharded_state_dict(model):
    """Per-rank checkpoint of a distributed model whose tables are too large to re-assemble: dense parameters (identical on every
    rank) + this rank's shard buffers and the routing needed to interpret them (row r of table t = shard[base[t] + r // world] on
    rank r % world).  Collective only through the flush of the lazy table semantics."""
    ctx = model._dist
    sh = ctx.sharded
    opt = getattr(model, "optim", None)
    if opt is not None and hasattr(opt, "flush"):
        opt.flush()
    table_keys = set("embedding_dict.%s.weight" % n for n in model.embedding_dict.keys()) | \
        set("linear_model.embedding_dict.%s.weight" % n for n in model.linear_model.embedding_dict.keys())
    dense = {k: v for k, v in torch.nn.Module.state_dict(model).items() if k not in table_keys}
    return {"dense": dense, "rank": ctx.rank, "world": ctx.world, "rows": list(sh.rows), "base": list(sh.base[ctx.rank]),
            "emb_shard": sh.emb.detach().clone(), "lin_shard": sh.lin.detach().clone()}
