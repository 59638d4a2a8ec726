"""Autograd-visible operators of the B200 hot path.  Every op enqueues hand-written sm_100a kernels from
libxdfm_sm100a.so on torch's current stream; torch only owns memory and the autograd tape.

Operators (reference code they replace):
  split_input      X[:, i:i+1].long() / dense column views      deepctr/models/basemodel.py:368-370, 377-378
  SparseGather     26x nn.Embedding + cat(dim=1)                deepctr/models/basemodel.py:354-380, xdeepfm.py:86
  LinearTerm       Linear.forward                               deepctr/models/basemodel.py:63-92
  CINFunction      CIN.forward                                  deepctr/layers/interaction.py:207-248
  LinearAct        nn.Linear -> activation                      deepctr/layers/core.py:120-134
  LogitHead        dnn_linear/cin_linear + sum + PredictionLayer deepctr/models/xdeepfm.py:88-105, layers/core.py:154-160
"""
import torch

from . import _native as N

_WS = {}
_WS_GENERATION = [0]      # bumped whenever a cached workspace is REPLACED: CUDA graphs captured before hold the old address

# optional per-operator device timing (bench.py / profiling): name -> [(start_event, end_event), ...]
TIMERS = None


class timed:
    """`with timed("cin_fwd"):` brackets the enclosed kernel launches with CUDA events on the current stream
    when TIMERS is a dict; otherwise it costs one attribute check."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if TIMERS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if TIMERS is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            TIMERS.setdefault(self.name, []).append((self.e0, e1))
        return False


def timer_totals():
    """ms per operator name (call after torch.cuda.synchronize())."""
    return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in (TIMERS or {}).items()}


def workspace(name, nbytes, device):
    """Cached scratch buffer (uint8) per (name, device); grows monotonically."""
    key = (name, device.index if device.index is not None else torch.cuda.current_device())
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _WS_GENERATION[0] += 1
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def workspace_generation():
    """Changes whenever a cached workspace was replaced by a larger one (graphs captured earlier must not be replayed)."""
    return _WS_GENERATION[0]


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _f32a(t):
    """contiguous fp32 at a 16-byte aligned address (128-bit loads): a view at an odd storage offset is copied."""
    t = _f32c(t)
    return t if t.data_ptr() % 16 == 0 else t.clone()


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("%s: the xdeepfm-b200 path runs on CUDA (sm_100a) only; got a %s tensor. "
                           "There is no CPU fallback -- construct the model with device='cuda:0'." % (what, t.device))


# ------------------------------------------------------------------------------------------------
# input split
# ------------------------------------------------------------------------------------------------
def split_input(X, sparse_cols, dense_cols):
    """X float32 [B, ncol] -> (ids int32 [B, m], dense float32 [B, nd])."""
    require_cuda(X, "split_input")
    X = _f32c(X)
    B, ncol = X.shape
    m, nd = len(sparse_cols), len(dense_cols)
    ids = torch.empty((B, m), dtype=torch.int32, device=X.device)
    dense = torch.empty((B, nd), dtype=torch.float32, device=X.device)
    N.check(N.lib().xdfm_split_input(N.ptr(X), B, ncol, N.i32_array(sparse_cols), m, N.i32_array(dense_cols), nd,
                                     N.ptr(ids), N.ptr(dense), N.stream_ptr()))
    return ids, dense


# ------------------------------------------------------------------------------------------------
# sparse plan + segment cache shared by the embedding and linear-term lookups of one step
# ------------------------------------------------------------------------------------------------
class SparsePlan:
    """Static description of a set of sparse features looked up together.

    feature f -> table index table_of[f]; tables have `rows[t]` rows; the scatter-add key space is the
    concatenation of the tables (row_off).  `sparse_grad` = True routes the backward result to `self.stash`
    (unique rows + summed gradients, consumed by the fused optimizer) instead of dense autograd gradients.
    """

    def __init__(self, table_of, rows, width):
        self.table_of = list(table_of)
        self.rows = list(rows)
        self.width = int(width)
        self.m = len(self.table_of)
        self.T = len(self.rows)
        if self.m > N.MAX_FIELDS:
            raise ValueError("at most %d sparse features are supported" % N.MAX_FIELDS)
        self.row_off = [0]
        for r in self.rows:
            self.row_off.append(self.row_off[-1] + int(r))
        self.vocab = [self.rows[t] for t in self.table_of]
        self.feat_row_off = [self.row_off[t] for t in self.table_of]
        self.sparse_grad = False
        self.stash = None
        self._c_vocab = N.i32_array(self.vocab)
        self._c_feat_off = N.i64_array(self.feat_row_off)
        self._c_row_off = N.i64_array(self.row_off)

    def signature(self):
        return (tuple(self.table_of), tuple(self.rows))

    def __reduce__(self):       # the ctypes arrays are rebuilt by the constructor (torch.save(model), copy.deepcopy)
        return (SparsePlan, (self.table_of, self.rows, self.width))


class SegmentCache:
    """Sort + run-length segments of one ids tensor, reused by every lookup that shares the key space.

    The key contains the device address of `ids`; the entry therefore keeps a strong reference to that tensor so the allocator
    cannot hand the address to another batch while the entry is alive (a freed-and-recycled allocation would otherwise look
    identical: same pointer, version 0, same shape)."""

    def __init__(self):
        self.key = None
        self.val = None
        self.ids = None

    def get(self, plan, ids):
        key = (ids.data_ptr(), ids._version, tuple(ids.shape), plan.signature())
        if self.key == key and self.ids is not None:
            return self.val
        B, m = ids.shape
        n = B * m
        dev = ids.device
        uniq = torch.empty(max(n, 1), dtype=torch.int32, device=dev)       # uint32 bit pattern
        seg_off = torch.empty(n + 1, dtype=torch.int32, device=dev)
        pos = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        nseg = torch.zeros(1, dtype=torch.int32, device=dev)
        nb = N.lib().xdfm_embed_bwd_workspace_bytes(n)
        ws = workspace("embed_bwd", nb, dev)
        N.check(N.lib().xdfm_embed_bwd_segments(N.ptr(ids), B, m, plan._c_feat_off, plan._c_vocab, plan.row_off[-1],
                                                N.ptr(ws), ws.numel(), N.ptr(uniq), N.ptr(seg_off), N.ptr(pos), N.ptr(nseg),
                                                N.stream_ptr()))
        self.key, self.val, self.ids = key, (uniq, seg_off, pos, nseg, n), ids
        return self.val

    def clear(self):
        self.key = self.val = self.ids = None


def segment_reduce(plan, seg, demb, dlin, width):
    """(gsum [n, width] or None, gsum_lin [n] or None) for the segments `seg`."""
    uniq, seg_off, pos, nseg, n = seg
    dev = uniq.device
    gsum = torch.empty((max(n, 1), width), dtype=torch.float32, device=dev) if demb is not None else None
    gsum_lin = torch.empty(max(n, 1), dtype=torch.float32, device=dev) if dlin is not None else None
    with timed("embed_scatter"):
        N.check(N.lib().xdfm_embed_bwd_reduce(N.ptr(demb), N.ptr(dlin), N.ptr(pos), N.ptr(seg_off), N.ptr(nseg), n, plan.m,
                                              width, N.ptr(gsum), N.ptr(gsum_lin), N.stream_ptr()))
    return gsum, gsum_lin


def scatter_dense(plan, seg, gsum, width, tables):
    """Dense per-table gradients (zeros + scatter of the segment sums)."""
    uniq, seg_off, pos, nseg, n = seg
    grads = [torch.zeros_like(t) for t in tables]
    N.check(N.lib().xdfm_embed_bwd_scatter_dense(N.ptr_array(grads), plan._c_row_off, plan.T, width, N.ptr(uniq), N.ptr(gsum),
                                                 N.ptr(nseg), n, N.stream_ptr()))
    return grads


class SparseGather(torch.autograd.Function):
    """out[b, f, :] = tables[table_of[f]][ids[b, f], :]  ->  [B, m, D]"""

    @staticmethod
    def forward(ctx, plan, cache, ids, *tables):
        require_cuda(ids, "SparseGather")
        B, m = ids.shape
        D = plan.width
        out = torch.empty((B, m, D), dtype=torch.float32, device=ids.device)
        per_feat = [tables[t] for t in plan.table_of]
        with timed("embed_gather"):
            N.check(N.lib().xdfm_embed_gather(N.ptr_array(per_feat), None, plan._c_vocab, N.ptr(ids), B, m, D, N.ptr(out),
                                              None, 0, None, None, N.stream_ptr()))
        ctx.plan, ctx.cache, ctx.ids, ctx.tables = plan, cache, ids, tables
        return out

    @staticmethod
    def backward(ctx, dout):
        plan = ctx.plan
        dout = _f32c(dout)
        seg = ctx.cache.get(plan, ctx.ids)
        gsum, _ = segment_reduce(plan, seg, dout, None, plan.width)
        if plan.sparse_grad:
            plan.stash = (seg, gsum)
            return (None, None, None) + tuple(None for _ in ctx.tables)
        grads = scatter_dense(plan, seg, gsum, plan.width, ctx.tables)
        return (None, None, None) + tuple(grads)


class LinearTerm(torch.autograd.Function):
    """lin[b] = sum_f lin_tables[table_of[f]][ids[b, f]] + dense[b, :] @ dense_w   ->  [B, 1]"""

    @staticmethod
    def forward(ctx, plan, cache, ids, dense, dense_w, *lin_tables):
        require_cuda(ids, "LinearTerm")
        B, m = ids.shape
        nd = 0 if dense is None or dense_w is None else dense.shape[1]
        out = torch.empty((B,), dtype=torch.float32, device=ids.device)
        per_feat = [lin_tables[t] for t in plan.table_of]
        N.check(N.lib().xdfm_embed_gather(None, N.ptr_array(per_feat) if m > 0 else None, plan._c_vocab, N.ptr(ids), B, m, 1, None,
                                          N.ptr(dense) if nd > 0 else None, nd, N.ptr(dense_w) if nd > 0 else None,
                                          N.ptr(out), N.stream_ptr()))
        ctx.plan, ctx.cache, ctx.ids, ctx.lin_tables = plan, cache, ids, lin_tables
        ctx.dense, ctx.nd = dense, nd
        ctx.dense_w_shape = None if dense_w is None else dense_w.shape
        return out.view(B, 1)

    @staticmethod
    def backward(ctx, dout):
        plan = ctx.plan
        dlin = _f32c(dout).reshape(-1)
        B = dlin.shape[0]
        d_dense_w = None
        if ctx.nd > 0:
            d_dense_w = torch.empty(ctx.nd, dtype=torch.float32, device=dlin.device)
            ws = workspace("wcolsum", N.lib().xdfm_wcolsum_workspace_bytes(ctx.nd), dlin.device)
            N.check(N.lib().xdfm_wcolsum(N.ptr(ctx.dense), B, ctx.nd, ctx.nd, N.ptr(dlin), N.ptr(d_dense_w), 0, N.ptr(ws),
                                         ws.numel(), N.stream_ptr()))
            d_dense_w = d_dense_w.view(ctx.dense_w_shape)
        table_grads = tuple(None for _ in ctx.lin_tables)
        if plan.m > 0:
            seg = ctx.cache.get(plan, ctx.ids)
            _, gsum_lin = segment_reduce(plan, seg, None, dlin, 1)
            if plan.sparse_grad:
                plan.stash = (seg, gsum_lin)
            else:
                table_grads = tuple(scatter_dense(plan, seg, gsum_lin, 1, ctx.lin_tables))
        return (None, None, None, None, d_dense_w) + table_grads


class BagLayout:
    """Slots -> fields map of one lookup with multi-value (VarLenSparseFeat) features (csrc/bag.cu).

    fields: list of (n_slots, mode, lencol): mode in 'single' | 'sum' | 'mean' | 'max' (the feature's `combiner`),
    lencol = column of the `lens` tensor holding the sequence length, or -1 for the id != 0 mask (inputs.py:141-155)."""

    def __init__(self, fields):
        self.fields = [tuple(f) for f in fields]
        self.F = len(fields)
        slot0, next_slot = [], 0
        for n, mode, _ in fields:
            if mode not in N.BAG:
                raise ValueError("parameter mode should in [sum, mean, max]")
            slot0.append(next_slot)
            next_slot += int(n)
        self.S = next_slot
        if self.S > N.MAX_FIELDS:
            raise ValueError("sparse features + sequence positions looked up together: %d, at most %d are supported" % (self.S, N.MAX_FIELDS))
        self.any_max = any(mode == "max" for _, mode, _ in fields)
        self.any_mean = any(mode == "mean" for _, mode, _ in fields)
        self.nlen = 1 + max([lc for _, _, lc in fields] + [-1])
        self._c_slot0 = N.i32_array(slot0)
        self._c_slen = N.i32_array([n for n, _, _ in fields])
        self._c_mode = N.i32_array([N.BAG[mode] for _, mode, _ in fields])
        self._c_lencol = N.i32_array([lc for _, _, lc in fields])

    def __reduce__(self):
        return (BagLayout, (self.fields,))


class BagPool(torch.autograd.Function):
    """slot tensor [B, S, D] -> field tensor [B, F, D]: fixed fields copied, sequence fields pooled under their mask."""

    @staticmethod
    def forward(ctx, lay, emb, ids, lens):
        require_cuda(emb, "BagPool")
        emb = _f32c(emb)
        B, S, D = emb.shape
        if S != lay.S or tuple(ids.shape) != (B, S) or ids.dtype != torch.int32:
            raise ValueError("BagPool: slot tensor %s / ids %s %s do not match the layout (%d slots)" % (
                tuple(emb.shape), tuple(ids.shape), ids.dtype, lay.S))
        if lay.nlen > 0 and (lens is None or lens.dtype != torch.int32 or tuple(lens.shape) != (B, lay.nlen)):
            raise ValueError("BagPool: lens must be int32 [B, %d]" % lay.nlen)
        ids = ids.contiguous()
        lens = lens.contiguous() if lay.nlen > 0 else None
        out = torch.empty((B, lay.F, D), dtype=torch.float32, device=emb.device)
        argmax = torch.empty((B, lay.F, D), dtype=torch.int32, device=emb.device) if lay.any_max else None
        den = torch.empty((B, lay.F), dtype=torch.float32, device=emb.device) if lay.any_mean else None
        with timed("bag_pool"):
            N.check(N.lib().xdfm_bag_pool_fwd(N.ptr(emb), N.ptr(ids), N.ptr(lens), lay.nlen, B, S, D, lay.F, lay._c_slot0, lay._c_slen,
                                              lay._c_mode, lay._c_lencol, N.ptr(out), N.ptr(argmax), N.ptr(den), N.stream_ptr()))
        ctx.lay, ctx.ids, ctx.lens, ctx.argmax, ctx.den, ctx.shape = lay, ids, lens, argmax, den, (B, S, D)
        return out

    @staticmethod
    def backward(ctx, dout):
        lay = ctx.lay
        B, S, D = ctx.shape
        dout = _f32c(dout)
        demb = torch.empty((B, S, D), dtype=torch.float32, device=dout.device)
        with timed("bag_pool"):
            N.check(N.lib().xdfm_bag_pool_bwd(N.ptr(dout), N.ptr(ctx.ids), N.ptr(ctx.lens), lay.nlen, N.ptr(ctx.argmax), N.ptr(ctx.den), B, S, D,
                                              lay.F,
                                              lay._c_slot0, lay._c_slen, lay._c_mode, lay._c_lencol, N.ptr(demb), N.stream_ptr()))
        return None, demb, None, None


# ------------------------------------------------------------------------------------------------
# CIN
# ------------------------------------------------------------------------------------------------
class CINConfig:
    def __init__(self, field_size, layer_size, split_half, activation, pool=True, impl="fp32"):
        self.m = int(field_size)
        self.layer_size = tuple(int(s) for s in layer_size)
        self.split_half = bool(split_half)
        self.act = N.ACT[activation]
        self.pool = bool(pool)
        self.impl = impl
        n = len(self.layer_size)
        self.Hp, self.direct_begin, self.n_next, self.col_off = [], [], [], []
        prev, col = self.m, 0
        for k, H in enumerate(self.layer_size):
            self.Hp.append(prev)
            last = k == n - 1
            if self.split_half:
                db = 0 if last else H // 2
                nxt = 0 if last else H // 2
            else:
                db = 0
                nxt = 0 if last else H
            self.direct_begin.append(db)
            self.n_next.append(nxt)
            self.col_off.append(col)
            col += H - db
            prev = H // 2 if self.split_half else H
        self.fm = col


class CINFunction(torch.autograd.Function):
    """CIN forward/backward; args: cfg, x0 [B,m,D], then W_0, b_0, W_1, b_1, ..."""

    @staticmethod
    def forward(ctx, cfg, x0, *wb):
        require_cuda(x0, "CIN")
        x0 = _f32c(x0)
        B, m, D = x0.shape
        dev = x0.device
        L = N.lib()
        out = torch.empty((B, cfg.fm) if cfg.pool else (B, cfg.fm, D), dtype=torch.float32, device=dev)
        ys = []
        xk, xk_stride = x0, m * D
        for k, H in enumerate(cfg.layer_size):
            W = _f32c(wb[2 * k]).view(H, -1)
            b = _f32c(wb[2 * k + 1])
            y = torch.empty((B, H, D), dtype=torch.float32, device=dev)
            with timed("cin_fwd"):
                N.check(L.xdfm_cin_fwd_f32(N.ptr(x0), N.ptr(xk), xk_stride, N.ptr(W), N.ptr(b), B, m, cfg.Hp[k], H, D, cfg.act,
                                           N.ptr(y), cfg.direct_begin[k], N.ptr(out) if cfg.pool else None,
                                           None if cfg.pool else N.ptr(out), cfg.fm, cfg.col_off[k], N.stream_ptr()))
            ys.append(y)
            xk, xk_stride = y, H * D
        ctx.cfg, ctx.x0, ctx.ys, ctx.wb = cfg, x0, ys, wb
        return out

    @staticmethod
    def backward(ctx, dout):
        cfg, x0, ys, wb = ctx.cfg, ctx.x0, ctx.ys, ctx.wb
        dout = _f32c(dout)
        B, m, D = x0.shape
        dev = x0.device
        L = N.lib()
        dx0 = torch.zeros_like(x0)
        grads = [None] * len(wb)
        dnext = None
        for k in range(len(cfg.layer_size) - 1, -1, -1):
            H, Hp = cfg.layer_size[k], cfg.Hp[k]
            y = ys[k]
            dy = torch.empty_like(y)
            N.check(L.xdfm_cin_dy(N.ptr(y), B, H, D, cfg.act, cfg.direct_begin[k], N.ptr(dout) if cfg.pool else None,
                                  None if cfg.pool else N.ptr(dout), cfg.fm, cfg.col_off[k], N.ptr(dnext), cfg.n_next[k],
                                  N.ptr(dy), N.stream_ptr()))
            xk, xk_stride = (x0, m * D) if k == 0 else (ys[k - 1], cfg.layer_size[k - 1] * D)
            W = _f32c(wb[2 * k]).view(H, -1)
            dW = torch.empty_like(W)
            db = torch.empty(H, dtype=torch.float32, device=dev)
            dxk = torch.empty((B, Hp, D), dtype=torch.float32, device=dev)
            nb = L.xdfm_cin_bwd_f32_workspace_bytes(B, m, Hp, H, D)
            ws = workspace("cin_bwd", nb, dev)
            with timed("cin_bwd"):
                N.check(L.xdfm_cin_bwd_f32(N.ptr(x0), N.ptr(xk), xk_stride, N.ptr(W), N.ptr(dy), B, m, Hp, H, D, N.ptr(dW), N.ptr(db),
                                           N.ptr(dxk), N.ptr(dx0), N.ptr(ws), ws.numel(), N.stream_ptr()))
            grads[2 * k] = dW.view(wb[2 * k].shape)
            grads[2 * k + 1] = db
            dnext = dxk
        dx0.add_(dnext)   # layer 0: X^{k-1} is X^0 itself
        return (None, dx0) + tuple(grads)


def _r8(v):
    return (v + 7) // 8 * 8


def _r16(v):
    return (v + 15) // 16 * 16


class CINFunctionTC(torch.autograd.Function):
    """CIN on the tensor cores (bf16 operands, fp32 accumulation): same signature as CINFunction.

    Activations live in the row layout of csrc/cin_tc.cu ([B*D, channels] bf16); the forward keeps x0t and every layer's yt for
    the backward, which runs per layer: dy (elementwise) -> dW (tcgen05, channel-major operands) -> dX (tcgen05)."""

    @staticmethod
    def forward(ctx, cfg, x0, *wb):
        require_cuda(x0, "CIN")
        x0 = _f32c(x0)
        B, m, D = x0.shape
        dev = x0.device
        L = N.lib()
        mP = _r8(m)
        out = torch.empty((B, cfg.fm) if cfg.pool else (B, cfg.fm, D), dtype=torch.float32, device=dev)
        x0t = torch.empty((B * D, mP), dtype=torch.bfloat16, device=dev)
        with timed("cin_layout"):
            N.check(L.xdfm_to_rows_bf16(N.ptr(x0), B, m, D, mP, N.ptr(x0t), N.stream_ptr()))
        yts = []
        xkt = x0t
        for k, H in enumerate(cfg.layer_size):
            W = _f32c(wb[2 * k]).view(H, -1)
            b = _f32c(wb[2 * k + 1])
            nw = L.xdfm_cin_tc_wprime_elems(m, cfg.Hp[k], H, D)
            if nw < 0:
                raise RuntimeError("libxdfm: " + L.xdfm_last_error().decode())
            wprime = workspace("cin_wprime", nw * 2, dev)
            yt = torch.empty((B * D, _r8(H)), dtype=torch.bfloat16, device=dev)
            with timed("cin_fwd"):
                N.check(L.xdfm_cin_fwd_tc(N.ptr(x0t), N.ptr(xkt), xkt.shape[1], N.ptr(W), N.ptr(b), N.ptr(wprime), B, m, cfg.Hp[k], H,
                                          D, cfg.act, N.ptr(yt), cfg.direct_begin[k], N.ptr(out) if cfg.pool else None,
                                          None if cfg.pool else N.ptr(out), cfg.fm, cfg.col_off[k], N.stream_ptr()))
            yts.append(yt)
            xkt = yt
        ctx.cfg, ctx.x0t, ctx.yts, ctx.wb, ctx.shape = cfg, x0t, yts, wb, (B, m, D)
        return out

    @staticmethod
    def backward(ctx, dout):
        cfg, x0t, yts, wb = ctx.cfg, ctx.x0t, ctx.yts, ctx.wb
        B, m, D = ctx.shape
        dout = _f32c(dout)
        dev = x0t.device
        L = N.lib()
        R, mP = B * D, _r8(m)
        st = N.stream_ptr()
        x0T = torch.empty((mP, R), dtype=torch.bfloat16, device=dev)
        with timed("cin_layout"):
            N.check(L.xdfm_rows_to_cols_bf16(N.ptr(x0t), mP, R, m, mP, N.ptr(x0T), st))
        n_layers = len(cfg.layer_size)
        dx0_parts = torch.empty((n_layers, 2, R, mP), dtype=torch.float32, device=dev)   # per layer: one dX0 plane per channel half
        grads = [None] * len(wb)
        dnext, dnext_pitch = None, 0
        dyt_ready = None        # dY rows of the layer about to be processed whose hidden half the dX kernel above has already written
        for k in range(len(cfg.layer_size) - 1, -1, -1):
            H, Hp = cfg.layer_size[k], cfg.Hp[k]
            Hs, H_pad, HpQ = _r8(H), _r16(H), _r16(Hp)
            yt = yts[k]
            xkt = x0t if k == 0 else yts[k - 1]
            dyt = torch.empty_like(yt) if dyt_ready is None else dyt_ready
            dyT = torch.empty((H_pad, R), dtype=torch.bfloat16, device=dev)
            reuse_x0T = k == 0 and HpQ == mP        # layer 0: X^{k-1} is X^0 itself and the channel-major copy already exists
            xkT = x0T if reuse_x0T else torch.empty((HpQ, R), dtype=torch.bfloat16, device=dev)
            with timed("cin_layout"):
                # (dnext = None with pitch -1: the hidden half of dyt is already there)
                db, db_ret = _grad_dst(wb[2 * k + 1], (H,))
                dbws = workspace("cin_db_part", L.xdfm_cin_dy_db_workspace_bytes(B, D, H_pad), dev)
                N.check(L.xdfm_cin_dy_rows_cols_db(N.ptr(yt), B, D, H, Hs, H_pad, cfg.direct_begin[k], N.ptr(dout) if cfg.pool else None,
                                                   None if cfg.pool else N.ptr(dout), cfg.fm, cfg.col_off[k], N.ptr(dnext),
                                                   -1 if dyt_ready is not None else dnext_pitch, cfg.n_next[k], cfg.act, N.ptr(dyt),
                                                   N.ptr(dyT), N.ptr(db), N.ptr(dbws), dbws.numel(), st))
                if not reuse_x0T:
                    N.check(L.xdfm_rows_to_cols_bf16(N.ptr(xkt), xkt.shape[1], R, Hp, HpQ, N.ptr(xkT), st))
            W = _f32c(wb[2 * k]).view(H, -1)
            dW, dW_ret = _grad_dst(wb[2 * k], W.shape)
            nb = L.xdfm_cin_bwd_dw_tc_workspace_bytes(B, m, Hp, H, D)
            if nb < 0:
                raise RuntimeError("libxdfm: " + L.xdfm_last_error().decode())
            ws = workspace("cin_dw_part", nb, dev)
            with timed("cin_bwd"):
                N.check(L.xdfm_cin_bwd_dw_tc(N.ptr(dyT), N.ptr(xkT), N.ptr(x0T), B, m, Hp, H, D, N.ptr(dW), None, N.ptr(ws),
                                             ws.numel(), st))            # (db came out of the dY kernel)
            nwt = L.xdfm_cin_bwd_dx_tc_wt_elems(m, Hp, H, D)
            wt = workspace("cin_wt", nwt * 2, dev)
            # the layer below gets its dY rows straight from this layer's dX kernel when its hidden half feeds only this layer
            # (split_half: channels [0, Hp) of y_{k-1}; direct-connect channels start at or after Hp) and the rows are wide enough
            fuse_dy = (k > 0 and cfg.n_next[k - 1] == Hp and cfg.direct_begin[k - 1] >= Hp and HpQ <= _r8(cfg.layer_size[k - 1])
                       and cfg.act in (N.ACT["relu"], N.ACT["linear"]))
            if fuse_dy:
                dyt_ready = torch.empty_like(yts[k - 1])
                dxk = None
                with timed("cin_bwd"):
                    N.check(L.xdfm_cin_bwd_dx_tc_dy(N.ptr(dyt), N.ptr(x0t), N.ptr(xkt), xkt.shape[1], N.ptr(W), N.ptr(wt), B, m, Hp, H, D,
                                                    None, N.ptr(dx0_parts[k]), N.ptr(dyt_ready), dyt_ready.shape[1], cfg.act, st))
            else:
                dyt_ready = None
                dxk = torch.empty((R, HpQ), dtype=torch.float32, device=dev)
                with timed("cin_bwd"):
                    N.check(L.xdfm_cin_bwd_dx_tc(N.ptr(dyt), N.ptr(x0t), N.ptr(xkt), xkt.shape[1], N.ptr(W), N.ptr(wt), B, m, Hp, H, D,
                                                 N.ptr(dxk), N.ptr(dx0_parts[k]), st))
            grads[2 * k] = dW_ret
            grads[2 * k + 1] = db_ret
            dnext, dnext_pitch = dxk, HpQ
        dx0 = torch.empty((B, m, D), dtype=torch.float32, device=dev)
        with timed("cin_layout"):
            # every layer's planes + layer 0's dXk (X^{k-1} is X^0 itself there), back in the reference layout
            N.check(L.xdfm_cin_dx0_finish(N.ptr(dx0_parts), 2 * n_layers, N.ptr(dnext), dnext_pitch, B, m, D, mP, N.ptr(dx0), st))
        return (None, dx0) + tuple(grads)


# ------------------------------------------------------------------------------------------------
# parameter gradients written in place (fused training step only)
# ------------------------------------------------------------------------------------------------
_DIRECT = {"on": False, "done": set()}


def direct_param_grads(on):
    """The fused training step zeroes one flat gradient buffer and binds every dense p.grad to a view of it.  While this is on, the
    backward kernels of the hot layers write a parameter's gradient straight into that view and hand autograd None, so no
    `p.grad += g` launch follows (one small add per parameter per step before).  Only the FIRST gradient of a parameter in a
    backward pass goes in place -- the view is still zero then; a parameter used twice gets its later contributions accumulated
    by autograd as usual."""
    _DIRECT["on"] = bool(on)
    _DIRECT["done"].clear()


def _grad_dst(param, shape=None):
    """-> (tensor the kernel writes d(param) into, value to return to autograd for it)."""
    shape = tuple(param.shape) if shape is None else tuple(shape)
    g = param.grad if (_DIRECT["on"] and isinstance(param, torch.nn.Parameter) and param.is_leaf and param.requires_grad) else None
    if g is not None and g.dtype == torch.float32 and g.is_cuda and g.is_contiguous() and g.shape == param.shape and \
            id(param) not in _DIRECT["done"]:
        _DIRECT["done"].add(id(param))
        return g.view(shape), None
    t = torch.empty(shape, dtype=torch.float32, device=param.device)
    return t, t.view(param.shape)


def cin_apply(cfg, x0, *wb):
    """Dispatch on the configured precision: 'fp32' = CUDA-core kernels (reference precision), 'bf16' = tcgen05 kernels."""
    if cfg.impl == "bf16":
        return CINFunctionTC.apply(cfg, x0, *wb)
    if cfg.impl != "fp32":
        raise ValueError("cin precision must be 'fp32' or 'bf16', got %r" % (cfg.impl,))
    return CINFunction.apply(cfg, x0, *wb)


# ------------------------------------------------------------------------------------------------
# dense layers
# ------------------------------------------------------------------------------------------------
def gemm(transA, transB, M, Nn, K, A, lda, Bm, ldb, C, ldc, bias=None, act=0, accumulate=False):
    L = N.lib()
    nb = L.xdfm_gemm_workspace_bytes(M, Nn, K)
    ws = workspace("gemm", nb, C.device) if nb > 0 else None
    with timed("gemm"):
        N.check(L.xdfm_gemm_f32(int(transA), int(transB), M, Nn, K, N.ptr(A), lda, N.ptr(Bm), ldb, N.ptr(C), ldc, N.ptr(bias), act,
                                int(accumulate), N.ptr(ws), 0 if ws is None else ws.numel(), N.stream_ptr()))


class LinearAct(torch.autograd.Function):
    """y = act(x @ W.T + b);  x [..., K], W [N, K], b [N] or None."""

    @staticmethod
    def forward(ctx, x, W, b, act):
        require_cuda(x, "LinearAct")
        shp = x.shape
        x2 = _f32c(x).reshape(-1, shp[-1])
        W = _f32c(W)
        Bn, K = x2.shape
        Nn = W.shape[0]
        y = torch.empty((Bn, Nn), dtype=torch.float32, device=x.device)
        gemm(0, 1, Bn, Nn, K, x2, K, W, K, y, Nn, bias=None if b is None else _f32c(b), act=act)
        ctx.save_for_backward(x2, W, y)
        ctx.act, ctx.has_bias, ctx.shp = act, b is not None, shp
        return y.view(*shp[:-1], Nn)

    @staticmethod
    def backward(ctx, dy):
        x2, W, y = ctx.saved_tensors
        Bn, K = x2.shape
        Nn = W.shape[0]
        dy = _f32c(dy).reshape(Bn, Nn)
        L = N.lib()
        if ctx.act != 0:
            dym = torch.empty_like(dy)
            N.check(L.xdfm_act_bwd(N.ptr(dy), N.ptr(y), N.ptr(dym), dy.numel(), ctx.act, N.stream_ptr()))
        else:
            dym = dy
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x2)
            gemm(0, 0, Bn, K, Nn, dym, Nn, W, K, dx, K)
            dx = dx.view(ctx.shp)
        if ctx.needs_input_grad[1]:
            dW = torch.empty_like(W)
            gemm(1, 0, Nn, K, Bn, dym, Nn, x2, K, dW, K)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(Nn, dtype=torch.float32, device=dy.device)
            ws = workspace("wcolsum", L.xdfm_wcolsum_workspace_bytes(Nn), dy.device)
            N.check(L.xdfm_wcolsum(N.ptr(dym), Bn, Nn, Nn, None, N.ptr(db), 0, N.ptr(ws), ws.numel(), N.stream_ptr()))
        return dx, dW, db, None


def cvt_bf16(src, transpose=False):
    """fp32 [R, C] -> bf16 [R, C8] (or, transposed, [C, R8]); pitch rounded up to 8 elements, padding zero."""
    R, C = src.shape
    rows, cols = (C, R) if transpose else (R, C)
    require_cuda(src, "cvt_bf16")
    if src.dtype != torch.float32 or (C > 1 and src.stride(1) != 1):
        src = _f32c(src)
    dst = torch.empty((rows, _r8(cols)), dtype=torch.bfloat16, device=src.device)
    N.check(N.lib().xdfm_cvt_bf16(ctypes_ptr_offset_any(src), R, C, max(src.stride(0), C), int(transpose), N.ptr(dst), dst.shape[1],
                                  N.stream_ptr()))
    return dst


def gemm_tc(A, Bm, M, Nn, K, bias=None, act=0, out=None):
    """fp32 [M, Nn] = act(A[M, K] . Bm[Nn, K]^T + bias) on tcgen05; A / Bm bf16 K-major (from cvt_bf16)."""
    L = N.lib()
    C = torch.empty((M, Nn), dtype=torch.float32, device=A.device) if out is None else out
    nb = L.xdfm_gemm_tc_workspace_bytes(M, Nn, K)
    ws = workspace("gemm_tc", nb, A.device)
    with timed("gemm"):
        N.check(L.xdfm_gemm_tc(M, Nn, K, N.ptr(A), A.shape[1], N.ptr(Bm), Bm.shape[1], N.ptr(C), Nn, N.ptr(bias), act, N.ptr(ws),
                               ws.numel(), N.stream_ptr()))
    return C


def cvt_bf16_both(src, want_rows=True, want_cols=True, y=None, act=0, want_colsum=False, colsum_out=None):
    """One pass over fp32 [R, C]: g = src * act'(y) -> (bf16 [R, C8] or None, bf16 [C, R8] or None, fp32 column sums [C] or None)."""
    R, C = src.shape
    require_cuda(src, "cvt_bf16_both")
    src = _f32c(src)
    dev = src.device
    L = N.lib()
    rows = torch.empty((R, _r8(C)), dtype=torch.bfloat16, device=dev) if want_rows else None
    cols = torch.empty((C, _r8(R)), dtype=torch.bfloat16, device=dev) if want_cols else None
    colsum = (torch.empty(C, dtype=torch.float32, device=dev) if colsum_out is None else colsum_out) if want_colsum else None
    ws = workspace("cvt_both", L.xdfm_cvt_bf16_both_workspace_bytes(R, C), dev) if want_colsum else None
    N.check(L.xdfm_cvt_bf16_both(N.ptr(src), N.ptr(y), int(act), R, C, C, N.ptr(rows), 0 if rows is None else rows.shape[1], N.ptr(cols),
                                 0 if cols is None else cols.shape[1], N.ptr(colsum), N.ptr(ws), 0 if ws is None else ws.numel(),
                                 N.stream_ptr()))
    return rows, cols, colsum


class LinearActTC(torch.autograd.Function):
    """y = act(x @ W.T + b) with bf16 operands on the tensor cores (fp32 accumulate, fp32 activations in HBM).

    Every bf16 operand comes out of ONE pass over its fp32 source (xdfm_cvt_bf16_both): the forward converts x and W once for both
    GEMM orientations (row-major for this GEMM, transposed for the backward's), the backward folds act'(y), both copies of dY and
    the bias gradient into one pass."""

    @staticmethod
    def forward(ctx, x, W, b, act):
        require_cuda(x, "LinearActTC")
        shp = x.shape
        x2 = _f32c(x).reshape(-1, shp[-1])
        W_in = W
        W = _f32c(W)
        Bn, K = x2.shape
        Nn = W.shape[0]
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        with timed("gemm_cvt"):
            xb, xT, _ = cvt_bf16_both(x2, True, need_dw)                 # [Bn, K8], [K, Bn8] (for dW = dY^T . x)
            wb, wT, _ = cvt_bf16_both(W, True, need_dx)                  # [Nn, K8], [K, Nn8] (for dX = dY . W)
        y = gemm_tc(xb, wb, Bn, Nn, K, None if b is None else _f32c(b), act)
        ctx.save_for_backward(xT, wT, y if act != 0 else None)
        ctx.act, ctx.has_bias, ctx.shp, ctx.dims = act, b is not None, shp, (Bn, K, Nn)
        ctx.params = (W_in, b)                  # the leaves themselves: their gradients may be written in place (direct_param_grads)
        return y.view(*shp[:-1], Nn)

    @staticmethod
    def backward(ctx, dy):
        xT, wT, y = ctx.saved_tensors
        Bn, K, Nn = ctx.dims
        dy = _f32c(dy).reshape(Bn, Nn)
        need_dx, need_dw, need_db = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        W_in, b_in = ctx.params
        db_dst = db_ret = None
        if need_db:
            db_dst, db_ret = _grad_dst(b_in, (Nn,))
        with timed("gemm_cvt"):
            dyb, dyT, _ = cvt_bf16_both(dy, need_dx, need_dw, y=y, act=ctx.act, want_colsum=need_db, colsum_out=db_dst)
        dx = dW_ret = None
        if need_dx:
            dx = gemm_tc(dyb, wT, Bn, K, Nn).view(ctx.shp)
        if need_dw:
            dW, dW_ret = _grad_dst(W_in, (Nn, K))
            gemm_tc(dyT, xT, Nn, K, Bn, out=dW)
        return dx, dW_ret, db_ret, None


SMALL_LINEAR_MAX = 32


class SmallLinear(torch.autograd.Function):
    """(y_0, ..) = act(x @ W_q.T + b) for up to three [N, K] weights sharing x, K, N <= 32 (csrc/smalllin.cu): the attention
    block's projections over B*L rows are HBM streaming, not GEMMs -- one pass over x yields Q, K and V; exact fp32."""

    @staticmethod
    def forward(ctx, x, b, act, *Ws):
        require_cuda(x, "SmallLinear")
        shp = x.shape
        x2 = _f32a(x).reshape(-1, shp[-1])
        Ws = [_f32c(W) for W in Ws]
        R, K = x2.shape
        Nn, nq = Ws[0].shape[0], len(Ws)
        ys = [torch.empty((R, Nn), dtype=torch.float32, device=x.device) for _ in range(nq)]
        wp = [N.ptr(W) for W in Ws] + [None] * (3 - nq)
        yp = [N.ptr(y) for y in ys] + [None] * (3 - nq)
        bias = None if b is None else _f32c(b)
        with timed("small_linear"):
            N.check(N.lib().xdfm_small_linear_fwd(N.ptr(x2), wp[0], wp[1], wp[2], N.ptr(bias), act, R, K, Nn, nq, yp[0], yp[1], yp[2],
                                                  N.stream_ptr()))
        ctx.save_for_backward(x2, *Ws, *(ys if act != 0 else []))
        ctx.act, ctx.has_bias, ctx.shp, ctx.nq = act, b is not None, shp, nq
        outs = tuple(y.view(*shp[:-1], Nn) for y in ys)
        return outs if nq > 1 else outs[0]

    @staticmethod
    def backward(ctx, *dys):
        saved = ctx.saved_tensors
        nq = ctx.nq
        x2, Ws = saved[0], saved[1:1 + nq]
        ys = saved[1 + nq:]
        R, K = x2.shape
        Nn = Ws[0].shape[0]
        L = N.lib()
        dev = x2.device
        dym = []
        for q in range(nq):
            dy = dys[q]
            dy = torch.zeros((R, Nn), dtype=torch.float32, device=dev) if dy is None else _f32a(dy).reshape(R, Nn)
            if ctx.act != 0:
                t = torch.empty_like(dy)
                N.check(L.xdfm_act_bwd(N.ptr(dy), N.ptr(ys[q]), N.ptr(t), dy.numel(), ctx.act, N.stream_ptr()))
                dy = t
            dym.append(dy)
        dp = [N.ptr(d) for d in dym] + [None] * (3 - nq)
        wp = [N.ptr(W) for W in Ws] + [None] * (3 - nq)
        dx = None
        with timed("small_linear"):
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x2)
                N.check(L.xdfm_small_linear_bwd_dx(dp[0], dp[1], dp[2], wp[0], wp[1], wp[2], R, K, Nn, nq, N.ptr(dx), N.stream_ptr()))
                dx = dx.view(ctx.shp)
            dWs = [torch.empty_like(W) for W in Ws]
            db = torch.empty(Nn, dtype=torch.float32, device=dev) if ctx.has_bias else None
            ws = workspace("small_linear_dw", L.xdfm_small_linear_bwd_dw_workspace_bytes(R, K, Nn, nq), dev)
            gp = [N.ptr(g) for g in dWs] + [None] * (3 - nq)
            N.check(L.xdfm_small_linear_bwd_dw(N.ptr(x2), dp[0], dp[1], dp[2], R, K, Nn, nq, gp[0], gp[1], gp[2], N.ptr(db), N.ptr(ws),
                                               N.stream_ptr()))
        return (dx, db, None) + tuple(dWs)


def small_linear_ok(x, *Ws):
    return all(W.shape[0] <= SMALL_LINEAR_MAX and W.shape[1] <= SMALL_LINEAR_MAX for W in Ws) and 1 <= len(Ws) <= 3


def linear_multi(x, Ws, precision="fp32"):
    """[x @ W.T for W in Ws] (bias-free, same shapes): one fused pass when the layers are narrow, else one GEMM each."""
    if small_linear_ok(x, *Ws) and len(set(tuple(W.shape) for W in Ws)) == 1:
        out = SmallLinear.apply(x, None, 0, *Ws)
        return list(out) if len(Ws) > 1 else [out]
    return [linear_act(x, W, precision=precision) for W in Ws]


def linear_act(x, W, b=None, activation=None, precision="fp32"):
    """precision 'fp32' = exact CUDA-core SGEMM, 'bf16' = tcgen05 GEMM (bf16 operands, fp32 accumulate); layers with K, N <= 32
    (attention projections, pooling MLP) stream through the narrow-layer kernels in exact fp32 whatever the precision."""
    if small_linear_ok(x, W):
        return SmallLinear.apply(x, b, N.ACT[activation], W)
    if precision == "bf16":
        return LinearActTC.apply(x, W, b, N.ACT[activation])
    return LinearAct.apply(x, W, b, N.ACT[activation])


class LogitHead(torch.autograd.Function):
    """y_pred [B,1] = sigmoid(lin + dnn_out @ w_dnn.T + cin_out @ w_cin.T + bias)  (binary) or the raw sum."""

    @staticmethod
    def forward(ctx, lin, cin_out, w_cin, dnn_out, w_dnn, bias, binary):
        ref = lin if lin is not None else (cin_out if cin_out is not None else dnn_out)
        require_cuda(ref, "LogitHead")
        B = ref.shape[0]
        lin_c = None if lin is None else _f32c(lin).reshape(-1)
        cin_c = None if cin_out is None else _f32c(cin_out)
        dnn_c = None if dnn_out is None else _f32c(dnn_out)
        wc = None if w_cin is None else _f32c(w_cin).reshape(-1)
        wd = None if w_dnn is None else _f32c(w_dnn).reshape(-1)
        bs = None if bias is None else _f32c(bias)
        fm = 0 if cin_c is None else cin_c.shape[1]
        hd = 0 if dnn_c is None else dnn_c.shape[1]
        y = torch.empty(B, dtype=torch.float32, device=ref.device)
        N.check(N.lib().xdfm_head_fwd(N.ptr(lin_c), N.ptr(cin_c), N.ptr(wc), fm, N.ptr(dnn_c), N.ptr(wd), hd, N.ptr(bs), B,
                                      int(binary), N.ptr(y), N.stream_ptr()))
        ctx.saved = (cin_c, wc, dnn_c, wd, y)
        ctx.meta = (B, fm, hd, int(binary), None if lin is None else lin.shape, None if w_cin is None else w_cin.shape,
                    None if w_dnn is None else w_dnn.shape, bias is not None)
        return y.view(B, 1)

    @staticmethod
    def backward(ctx, dy):
        cin_c, wc, dnn_c, wd, y = ctx.saved
        B, fm, hd, binary, lin_shape, wc_shape, wd_shape, has_bias = ctx.meta
        dev = y.device
        L = N.lib()
        dy = _f32c(dy).reshape(-1)
        dlogit = torch.empty(B, dtype=torch.float32, device=dev)
        d_cin = torch.empty_like(cin_c) if cin_c is not None else None
        d_dnn = torch.empty_like(dnn_c) if dnn_c is not None else None
        # one pass: dlogit, the two output gradients and (d_w_cin | d_w_dnn | d_bias) as slices of one buffer
        fm_, hd_ = (fm if cin_c is not None else 0), (hd if dnn_c is not None else 0)
        d_w = torch.empty(fm_ + hd_ + 1, dtype=torch.float32, device=dev)
        ws = workspace("head_bwd", L.xdfm_head_bwd_fused_workspace_bytes(B, fm_, hd_), dev)
        N.check(L.xdfm_head_bwd_fused(N.ptr(dy), N.ptr(y), B, binary, N.ptr(cin_c), N.ptr(wc), fm_, N.ptr(dnn_c), N.ptr(wd), hd_,
                                      N.ptr(dlogit), N.ptr(d_cin), N.ptr(d_dnn), N.ptr(d_w), N.ptr(ws), ws.numel(), N.stream_ptr()))
        d_wc = d_w[:fm_].view(wc_shape) if cin_c is not None else None
        d_wd = d_w[fm_:fm_ + hd_].view(wd_shape) if dnn_c is not None else None
        d_bias = d_w[fm_ + hd_:] if has_bias else None
        d_lin = dlogit.view(lin_shape) if lin_shape is not None else None
        return d_lin, d_cin, d_wc, d_dnn, d_wd, d_bias, None


def bce_sum(y_pred, labels, loss_accum=None, scale=1.0, want_grad=True):
    """Fused F.binary_cross_entropy(y_pred, y, reduction='sum').
    Adds the loss into `loss_accum` (float64 [1] device tensor) and returns dL/dy_pred (or None)."""
    require_cuda(y_pred, "bce_sum")
    p = _f32c(y_pred).reshape(-1)
    y = _f32c(labels).reshape(-1)
    B = p.shape[0]
    if loss_accum is None:
        loss_accum = torch.zeros(1, dtype=torch.float64, device=p.device)
    g = torch.empty_like(p) if want_grad else None
    N.check(N.lib().xdfm_bce_sum(N.ptr(p), N.ptr(y), B, float(scale), None, N.ptr(g), N.ptr(loss_accum), N.stream_ptr()))
    return loss_accum, g


# ------------------------------------------------------------------------------------------------
# field self-attention block over the CIN feature maps (deepctr/layers/cin_attention.py)
# ------------------------------------------------------------------------------------------------
_DROPOUT_STATE = {}


def dropout_state(device):
    """Device-resident uint64 counter [1] feeding the in-kernel dropout hashes (one per device).  Seeded from torch's CUDA generator
    on first use (torch.manual_seed makes runs reproducible); every consumer clones the current value and then advances the
    counter with a device-side add, so the sequence survives CUDA-graph replay (a host-drawn seed would be baked into the graph)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _DROPOUT_STATE.get(key)
    if st is None:
        base = int(torch.randint(0, 2 ** 62, (1,), device=device).item())
        st = _DROPOUT_STATE[key] = torch.tensor([base], dtype=torch.int64, device=device)
    return st


def reset_dropout_state(seed=None):
    """Forget the per-device dropout counters (the next use re-seeds from torch's generator) or set them to `seed`."""
    if seed is None:
        _DROPOUT_STATE.clear()
    else:
        for st in _DROPOUT_STATE.values():
            st.fill_(int(seed))


class MHSACore(torch.autograd.Function):
    """o = dropout(softmax(q k^T / sqrt(head_dim))) v per head; q, k, v [B, L, E].  Scores never leave the SM; the backward
    recomputes the probabilities from the saved log-sum-exp -- and, with dropout, the keep mask from the saved seed (reference:
    cin_attention.py:73-95 materialises [B, h, L, L] scores, probabilities and the dropout mask)."""

    @staticmethod
    def forward(ctx, q, k, v, heads, dropout_p=0.0):
        require_cuda(q, "MHSACore")
        q, k, v = _f32c(q), _f32c(k), _f32c(v)
        B, L, E = q.shape
        o = torch.empty_like(q)
        lse = torch.empty((B, heads, L), dtype=torch.float32, device=q.device)
        seed = None
        with timed("mhsa"):
            if dropout_p > 0.0:
                st = dropout_state(q.device)
                seed = st.clone()
                st.add_(0x9E3779B97F4A7C15 & (2 ** 62 - 1))      # next call: a different stream
                N.check(N.lib().xdfm_mhsa_fwd_dropout(N.ptr(q), N.ptr(k), N.ptr(v), B, L, E, heads, float(dropout_p), N.ptr(seed), N.ptr(o),
                                                      N.ptr(lse), N.stream_ptr()))
            else:
                N.check(N.lib().xdfm_mhsa_fwd(N.ptr(q), N.ptr(k), N.ptr(v), B, L, E, heads, N.ptr(o), N.ptr(lse), N.stream_ptr()))
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.heads, ctx.dropout_p, ctx.seed = heads, float(dropout_p), seed
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        B, L, E = q.shape
        do = _f32c(do)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        with timed("mhsa"):
            if ctx.dropout_p > 0.0:
                N.check(N.lib().xdfm_mhsa_bwd_dropout(N.ptr(q), N.ptr(k), N.ptr(v), N.ptr(o), N.ptr(lse), N.ptr(do), B, L, E, ctx.heads,
                                                      ctx.dropout_p, N.ptr(ctx.seed), N.ptr(dq), N.ptr(dk), N.ptr(dv), N.stream_ptr()))
            else:
                N.check(N.lib().xdfm_mhsa_bwd(N.ptr(q), N.ptr(k), N.ptr(v), N.ptr(o), N.ptr(lse), N.ptr(do), B, L, E, ctx.heads, N.ptr(dq),
                                              N.ptr(dk), N.ptr(dv), N.stream_ptr()))
        return dq, dk, dv, None, None


class AddLayerNorm(torch.autograd.Function):
    """y = LayerNorm_E(a + r) * gamma + beta (r and the normalisation are optional): residual + nn.LayerNorm of
    cin_attention.py:305-311 in one pass."""

    @staticmethod
    def forward(ctx, a, r, gamma, beta, eps, normalize):
        require_cuda(a, "AddLayerNorm")
        a = _f32c(a)
        r = None if r is None else _f32c(r)
        E = a.shape[-1]
        rows = a.numel() // E
        y = torch.empty_like(a)
        mean = torch.empty(rows, dtype=torch.float32, device=a.device) if normalize else None
        rstd = torch.empty(rows, dtype=torch.float32, device=a.device) if normalize else None
        g = None if gamma is None else _f32c(gamma)
        bt = None if beta is None else _f32c(beta)
        N.check(N.lib().xdfm_add_ln_fwd(N.ptr(a), N.ptr(r), N.ptr(g), N.ptr(bt), rows, E, float(eps), int(bool(normalize)), N.ptr(y),
                                        N.ptr(mean), N.ptr(rstd), N.stream_ptr()))
        ctx.saved = (a, r, g, mean, rstd)
        ctx.normalize, ctx.has_r = bool(normalize), r is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        a, r, g, mean, rstd = ctx.saved
        dy = _f32c(dy)
        if not ctx.normalize:
            return dy, (dy if ctx.has_r else None), None, None, None, None
        E = a.shape[-1]
        rows = a.numel() // E
        L = N.lib()
        dx = torch.empty_like(a)
        nblk = L.xdfm_add_ln_bwd_blocks(rows)
        partial = torch.empty((nblk, 2 * E), dtype=torch.float32, device=a.device)
        N.check(L.xdfm_add_ln_bwd(N.ptr(dy), N.ptr(a), N.ptr(r), N.ptr(g), N.ptr(mean), N.ptr(rstd), rows, E, N.ptr(dx), N.ptr(partial),
                                  N.stream_ptr()))
        dgb = torch.empty(2 * E, dtype=torch.float32, device=a.device)
        ws = workspace("wcolsum", L.xdfm_wcolsum_workspace_bytes(2 * E), a.device)
        N.check(L.xdfm_wcolsum(N.ptr(partial), nblk, 2 * E, 2 * E, None, N.ptr(dgb), 0, N.ptr(ws), ws.numel(), N.stream_ptr()))
        return dx, (dx if ctx.has_r else None), dgb[:E], dgb[E:], None, None


class AttnPool(torch.autograd.Function):
    """out [B, E] = sum_l softmax_L(score)[b, l] * x[b, l, :]  (cin_attention.py:138-142)."""

    @staticmethod
    def forward(ctx, score, x):
        require_cuda(x, "AttnPool")
        x = _f32c(x)
        B, Lq, E = x.shape
        s = _f32c(score).reshape(B, Lq)
        attn = torch.empty((B, Lq), dtype=torch.float32, device=x.device)
        out = torch.empty((B, E), dtype=torch.float32, device=x.device)
        N.check(N.lib().xdfm_attn_pool_fwd(N.ptr(s), N.ptr(x), B, Lq, E, N.ptr(attn), N.ptr(out), N.stream_ptr()))
        ctx.save_for_backward(attn, x)
        ctx.score_shape = score.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        attn, x = ctx.saved_tensors
        B, Lq, E = x.shape
        dout = _f32c(dout)
        dscore = torch.empty_like(attn)
        dx = torch.empty_like(x)
        N.check(N.lib().xdfm_attn_pool_bwd(N.ptr(dout), N.ptr(attn), N.ptr(x), B, Lq, E, N.ptr(dscore), N.ptr(dx), N.stream_ptr()))
        return dscore.view(ctx.score_shape), dx


# ------------------------------------------------------------------------------------------------
# xDeepFM Pro: SFG reconstruction losses (deepctr/xdeepfm_pro/sfg_decoder.py:266-309)
# ------------------------------------------------------------------------------------------------
def sfg_row_weights(labels, positive_only):
    """[B] weights mask / num_positive (sfg_decoder.py:266-273); stays on the device."""
    require_cuda(labels, "sfg_row_weights")
    y = _f32c(labels).reshape(-1)
    w = torch.empty_like(y)
    N.check(N.lib().xdfm_sfg_row_weights(N.ptr(y), y.shape[0], int(bool(positive_only)), N.ptr(w), N.stream_ptr()))
    return w


def _sum1(row_loss):
    out = torch.empty(1, dtype=torch.float32, device=row_loss.device)
    n = row_loss.shape[0]
    if n == 0:
        return out.zero_()
    ws = workspace("wcolsum", N.lib().xdfm_wcolsum_workspace_bytes(1), row_loss.device)
    N.check(N.lib().xdfm_wcolsum(N.ptr(row_loss), n, 1, 1, None, N.ptr(out), 0, N.ptr(ws), ws.numel(), N.stream_ptr()))
    return out


class MaskedCE(torch.autograd.Function):
    """sum_r row_w[r] * cross_entropy(logits[r], ids[r, col])  ->  [1].  Forward and gradient in one pass per row."""

    @staticmethod
    def forward(ctx, logits, ids, col, row_w):
        require_cuda(logits, "MaskedCE")
        logits = _f32c(logits)
        R, V = logits.shape
        row_loss = torch.empty(R, dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        tgt = ids[:, col:]           # element (r, 0) at stride ids.shape[1]
        N.check(N.lib().xdfm_masked_ce(N.ptr(logits), ctypes_ptr_offset(ids, col), ids.shape[1], N.ptr(row_w), R, V, N.ptr(row_loss),
                                       N.ptr(dlogits), N.stream_ptr()))
        del tgt
        ctx.save_for_backward(dlogits)
        return _sum1(row_loss)

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g.reshape(1, 1), None, None, None


class MaskedMSE(torch.autograd.Function):
    """sum_r row_w[r] * mean_j (pred[r, j] - target[r, j])^2  ->  [1]."""

    @staticmethod
    def forward(ctx, pred, target, row_w):
        require_cuda(pred, "MaskedMSE")
        pred, target = _f32c(pred), _f32c(target)
        R, nd = pred.shape
        row_loss = torch.empty(R, dtype=torch.float32, device=pred.device)
        dpred = torch.empty_like(pred)
        N.check(N.lib().xdfm_masked_mse(N.ptr(pred), N.ptr(target), N.ptr(row_w), R, nd, N.ptr(row_loss), N.ptr(dpred), N.stream_ptr()))
        ctx.save_for_backward(dpred)
        return _sum1(row_loss)

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return dpred * g.reshape(1, 1), None, None


# ------------------------------------------------------------------------------------------------
# xDeepFM Pro: AutoDis dense-feature encoder (deepctr/xdeepfm_pro/autodis.py:100-127)
# ------------------------------------------------------------------------------------------------
class AutoDis(torch.autograd.Function):
    """out [B, nd*E]: per feature Linear(1,nb) -> LeakyReLU(0.2) -> Linear(nb,nb) -> softmax(. / temp) @ meta[f], one launch.

    Inputs: x [B, nd], meta [nd, nb, E], temp [nd], then the reference's per-feature projector tensors in module order
    (w1_f [nb,1], b1_f [nb], W2_f [nb,nb], b2_f [nb]) * nd.  They are packed per call (4 small stacks); the backward returns views
    of one packed gradient buffer, no gradient w.r.t. x (dense inputs are data)."""

    @staticmethod
    def forward(ctx, x, meta, temp, *proj):
        require_cuda(x, "AutoDis")
        x = _f32c(x)
        B, nd = x.shape
        _, nb, E = meta.shape
        if len(proj) != 4 * nd:
            raise ValueError("AutoDis: expected %d projector tensors, got %d" % (4 * nd, len(proj)))
        w1 = torch.stack([proj[4 * f].reshape(nb) for f in range(nd)])
        b1 = torch.stack([proj[4 * f + 1] for f in range(nd)])
        W2 = torch.stack([proj[4 * f + 2] for f in range(nd)])
        b2 = torch.stack([proj[4 * f + 3] for f in range(nd)])
        meta_c, temp_c = _f32c(meta), _f32c(temp)
        out = torch.empty((B, nd * E), dtype=torch.float32, device=x.device)
        with timed("autodis"):
            N.check(N.lib().xdfm_autodis_fwd(N.ptr(x), N.ptr(w1), N.ptr(b1), N.ptr(W2), N.ptr(b2), N.ptr(meta_c), N.ptr(temp_c), B, nd, nb,
                                             E, N.ptr(out), N.stream_ptr()))
        ctx.save_for_backward(x, w1, b1, W2, b2, meta_c, temp_c)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w1, b1, W2, b2, meta, temp = ctx.saved_tensors
        B, nd = x.shape
        _, nb, E = meta.shape
        L = N.lib()
        n_par = L.xdfm_autodis_param_count(nb, E)
        gpack = torch.empty((nd, n_par), dtype=torch.float32, device=x.device)
        ws = workspace("autodis_bwd", L.xdfm_autodis_bwd_workspace_bytes(B, nd, nb, E), x.device)
        dout = _f32c(dout)
        with timed("autodis"):
            N.check(L.xdfm_autodis_bwd(N.ptr(x), N.ptr(w1), N.ptr(b1), N.ptr(W2), N.ptr(b2), N.ptr(meta), N.ptr(temp), N.ptr(dout), B, nd,
                                       nb, E, N.ptr(gpack), N.ptr(ws), N.stream_ptr()))
        o = 0
        g_meta = gpack[:, o:o + nb * E].reshape(nd, nb, E); o += nb * E
        g_W2 = gpack[:, o:o + nb * nb]; o += nb * nb
        g_b2 = gpack[:, o:o + nb]; o += nb
        g_w1 = gpack[:, o:o + nb]; o += nb
        g_b1 = gpack[:, o:o + nb]; o += nb
        g_temp = gpack[:, o]
        grads = []
        for f in range(nd):
            grads += [g_w1[f].reshape(nb, 1), g_b1[f], g_W2[f].reshape(nb, nb), g_b2[f]]
        return (None, g_meta, g_temp) + tuple(grads)


def ctypes_ptr_offset_any(t):
    """Device pointer of a (possibly row-strided) CUDA tensor."""
    import ctypes
    return ctypes.c_void_p(t.data_ptr())


def ctypes_ptr_offset(t, elem_offset):
    """Device pointer of element `elem_offset` of a contiguous CUDA tensor."""
    import ctypes
    require_cuda(t, "pointer")
    if not t.is_contiguous():
        raise RuntimeError("xdeepfm-b200 ops need contiguous tensors")
    return ctypes.c_void_p(t.data_ptr() + int(elem_offset) * t.element_size())
