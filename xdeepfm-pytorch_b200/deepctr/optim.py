"""Fused optimizers (replace torch.optim.{SGD,Adam,Adagrad,RMSprop} chosen by name in the reference:
deepctr/models/basemodel.py:447-461) and the L2 regulariser (basemodel.py:412-428).

One flat fp32 buffer holds every dense parameter (CIN, DNN, heads, attention ...): a single kernel updates all of them.
Embedding tables (and the [V,1] first-order tables) are updated in place by two kernels per table set: the rows the batch
touched (unique rows + deterministic segment sums from ops.SparseGather/LinearTerm) and -- to keep the reference's dense
semantics, where EVERY row moves every step because of L2 and optimizer momentum -- a streaming pass over all other rows.
"""
import torch

from . import _native as N
from . import ops

_TORCH_DEFAULTS = {
    "sgd": dict(lr=0.01),
    "adam": dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8),
    "adagrad": dict(lr=0.01, lr_decay=0.0, eps=1e-10),
    "rmsprop": dict(lr=0.01, alpha=0.99, eps=1e-8),
}
_N_STATES = {"sgd": 0, "adam": 2, "adagrad": 1, "rmsprop": 1}


class TableSet:
    """One group of tables sharing a SparsePlan (embedding tables of width D, or first-order tables of width 1)."""

    def __init__(self, plan, params, l2):
        self.plan, self.params, self.l2 = plan, list(params), float(l2)
        self.s1 = self.s2 = None
        self.bitmap = None
        self.last = None            # int32 [total_rows]: step up to which each row is current (lazy dense semantics)

    def __getstate__(self):         # moments travel through FusedOptimizer.state_dict(); scratch is rebuilt by prepare()
        d = dict(self.__dict__)
        d["s1"] = d["s2"] = d["bitmap"] = d["last"] = None
        return d


class FusedOptimizer(torch.optim.Optimizer):
    """torch.optim.Optimizer-compatible (param_groups / state_dict) front end of the fused kernels."""

    def __init__(self, kind, dense_named_params, table_sets, l2_of_param, dist_ctx=None, l2_sharded=(0.0, 0.0), **overrides):
        if kind not in _TORCH_DEFAULTS:
            raise NotImplementedError(kind)
        defaults = dict(_TORCH_DEFAULTS[kind])
        defaults.update(overrides)
        self.kind = kind
        self.dense_named = list(dense_named_params)
        self.table_sets = list(table_sets)
        all_params = [p for _, p in self.dense_named] + [p for ts in self.table_sets for p in ts.params]
        super().__init__(all_params, defaults)
        self._l2_of_param = l2_of_param          # id(param) -> l2 strength (sum over registrations)
        self._flat = None
        self.sparse_embedding_update = False     # True: only rows touched by the batch are updated (NOT reference semantics)
        self.reg_accum = None                    # float64 [1] device: sum of l2*w^2 seen by the last step(s)
        self.steps = 0
        # Lazy form of the reference's dense table semantics: rows the batch does not touch are not streamed every step; their
        # (deterministic) L2 / momentum updates are replayed bit-for-bit when the row is next looked up or at flush() time
        # (csrc/optim.cu).  False = stream every table every step (first implementation; kept for A/B tests).
        self.lazy_tables = True
        # The replay cost of a row grows with the number of steps it is behind, so the backlog is bounded: every `flush_interval`
        # steps all rows are settled (one streaming pass; amortised cost ~ the replay ALU work, independent of the interval).
        self.flush_interval = 64
        self._last_flush_step = 0
        self._hist = None
        self._hist_cap = 1 << 16
        self._hist_base = 0
        self._dirty = False
        # multi-GPU (deepctr.distributed): dense gradients are all-reduced, the row-sharded tables are updated by their owner
        self.dist_ctx = dist_ctx
        self.l2_sharded = (float(l2_sharded[0]), float(l2_sharded[1]))     # (embedding tables, first-order tables)
        self.reg_accum_shard = None              # float64 [1]: this rank's share of the table regulariser

    # -------------------------------------------------------------------------------------------
    def _cfg(self, l2):
        g = self.param_groups[0]
        c = N.OptCfg()
        c.kind = N.OPT[self.kind]
        c.lr = float(g["lr"])
        b1, b2 = g.get("betas", (0.9, 0.999))
        c.beta1, c.beta2 = float(b1), float(b2)
        c.eps = float(g.get("eps", 1e-8))
        c.alpha = float(g.get("alpha", 0.99))
        c.lr_decay = float(g.get("lr_decay", 0.0))
        c.l2 = float(l2)
        return c

    def _dense_params(self):
        return [p for _, p in self.dense_named]

    def _flat_valid(self):
        if self._flat is None:
            return False
        w = self._flat["w"]
        off = 0
        for p in self._dense_params():
            if p.data_ptr() != w.data_ptr() + off * 4 or p.device != w.device:
                return False
            off += p.numel()
        return True

    def prepare(self):
        """(Re)build the flat views; called lazily and whenever the parameters were moved (.to(device))."""
        params = self._dense_params()
        if self._flat_valid():
            return
        dev = params[0].device if params else self.table_sets[0].params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedOptimizer needs the model on a CUDA device (no CPU fallback); got %s" % dev)
        n = sum(p.numel() for p in params)
        old = self._flat
        w = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
        g = torch.zeros(max(n, 1), dtype=torch.float32, device=dev)
        l2vec = torch.zeros(max(n, 1), dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            k = p.numel()
            w[off:off + k].copy_(p.data.reshape(-1))
            p.data = w[off:off + k].view(p.shape)
            if p.grad is not None:      # a backward ran before the flat views existed (user loop / generic fit path): keep its result
                g[off:off + k].copy_(p.grad.reshape(-1))
            p.grad = g[off:off + k].view(p.shape)
            l2vec[off:off + k] = self._l2_of_param.get(id(p), 0.0)
            off += k
        ns = _N_STATES[self.kind]
        s1 = torch.zeros_like(w) if ns >= 1 else None
        s2 = torch.zeros_like(w) if ns >= 2 else None
        if old is not None and old["w"].numel() == w.numel():
            if s1 is not None and old["s1"] is not None:
                s1.copy_(old["s1"])
            if s2 is not None and old["s2"] is not None:
                s2.copy_(old["s2"])
        opt_dev = torch.zeros(8, dtype=torch.float32, device=dev)
        if old is not None:
            opt_dev.copy_(old["opt_dev"])
        self._flat = dict(w=w, g=g, s1=s1, s2=s2, l2vec=l2vec, n=n, opt_dev=opt_dev, any_l2=bool((l2vec != 0).any().item()))
        self.reg_accum = torch.zeros(1, dtype=torch.float64, device=dev)
        self.reg_accum_shard = torch.zeros(1, dtype=torch.float64, device=dev)
        if self.dist_ctx is not None:
            self.dist_ctx.sharded.set_states(ns)        # moments live in the exported shard buffer (peers replay stale rows)
        for ts in self.table_sets:
            if ts.s1 is None or ts.s1[0].device != dev:
                ts.s1 = [torch.zeros_like(p.data) for p in ts.params] if ns >= 1 else None
                ts.s2 = [torch.zeros_like(p.data) for p in ts.params] if ns >= 2 else None
                ts.bitmap = torch.zeros((ts.plan.row_off[-1] + 31) // 32 + 1, dtype=torch.int32, device=dev)
                ts.last = torch.zeros(max(ts.plan.row_off[-1], 1), dtype=torch.int32, device=dev)
        if self._hist is None or self._hist.device != dev:
            self._hist = torch.zeros(self._hist_cap * 4, dtype=torch.float32, device=dev)
        self._publish_lazy()
        pending = self.__dict__.pop("_pending_state", None)
        if pending is not None:
            self._load_fused_state(pending)

    # -------------------------------------------------------------------------------------------
    # checkpointing (the reference saves no optimizer state; this adds it).  state_dict() settles every postponed
    # row update first, so the saved moments are those of a dense pass and a resumed run continues bit-identically.
    # -------------------------------------------------------------------------------------------
    def state_dict(self):
        """{'fused': {kind, steps, hyper-parameters, opt_dev, dense moments by parameter NAME, table moments by table-set / table
        index}}.  Row-sharded tables: each rank's dict holds the moments of its own shard (save one file per rank)."""
        hp = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        out = {"kind": self.kind, "steps": int(self.steps), "hyper": hp, "lazy_tables": bool(self.lazy_tables),
               "sparse_embedding_update": bool(self.sparse_embedding_update), "flush_interval": int(self.flush_interval)}
        if self._flat is None:
            return {"fused": out}
        self.flush()
        f = self._flat
        out["opt_dev"] = f["opt_dev"].detach().clone()
        dense, off = {}, 0
        for name, p in self.dense_named:
            k = p.numel()
            dense[name] = tuple(None if s is None else s[off:off + k].detach().clone().view(p.shape) for s in (f["s1"], f["s2"]))
            off += k
        out["dense"] = dense
        out["tables"] = [{"s1": None if ts.s1 is None else [t.detach().clone() for t in ts.s1],
                          "s2": None if ts.s2 is None else [t.detach().clone() for t in ts.s2]} for ts in self.table_sets]
        if self.dist_ctx is not None:
            out["shard"] = self.dist_ctx.sharded.optimizer_state()
        return {"fused": out}

    def load_state_dict(self, state_dict):
        st = state_dict["fused"]
        if st["kind"] != self.kind:
            raise ValueError("optimizer state of kind %r cannot be loaded into a %r optimizer" % (st["kind"], self.kind))
        for k, v in st["hyper"].items():
            self.param_groups[0][k] = v
        self.lazy_tables = st.get("lazy_tables", self.lazy_tables)
        self.sparse_embedding_update = st.get("sparse_embedding_update", self.sparse_embedding_update)
        self.flush_interval = st.get("flush_interval", self.flush_interval)
        if "opt_dev" not in st:          # saved before the first step
            self.steps = int(st["steps"])
            return
        if self._flat is None or not self._flat_valid():
            self._pending_state = st     # applied at the end of prepare()
            self.prepare()
        else:
            self._load_fused_state(st)

    def _load_fused_state(self, st):
        self.flush()
        f = self._flat
        f["opt_dev"].copy_(st["opt_dev"])
        off = 0
        for name, p in self.dense_named:
            k = p.numel()
            if name not in st["dense"]:
                raise KeyError("optimizer state has no entry for parameter %r" % name)
            for buf, src in zip((f["s1"], f["s2"]), st["dense"][name]):
                if buf is not None and src is not None:
                    buf[off:off + k].copy_(src.reshape(-1))
            off += k
        if len(st["tables"]) != len(self.table_sets):
            raise ValueError("optimizer state holds %d table sets, the model has %d" % (len(st["tables"]), len(self.table_sets)))
        self.steps = int(st["steps"])
        for ts, saved in zip(self.table_sets, st["tables"]):
            for mine, theirs in ((ts.s1, saved["s1"]), (ts.s2, saved["s2"])):
                if mine is not None and theirs is not None:
                    for a, b in zip(mine, theirs):
                        a.copy_(b)
            if ts.last is not None:
                ts.last.fill_(self.steps)          # every row is current at the saved step (state_dict() flushed)
        if self.dist_ctx is not None and "shard" in st:
            self.dist_ctx.sharded.load_optimizer_state(st["shard"], self.steps)
        self._hist_base = self.steps                # history slots are (step - base): a fresh window starts at the resumed step
        self._last_flush_step = self.steps
        self._dirty = False
        self._publish_lazy()

    def __getstate__(self):
        """torch.save(model) (ModelCheckpoint's default): torch.optim.Optimizer pickles defaults / state / param_groups only; keep
        this class's configuration and the moments (through state_dict()), drop device scratch and the process-group context."""
        if self.dist_ctx is not None:
            raise RuntimeError("the optimizer of a distributed model cannot be pickled: save optimizer.state_dict() per rank")
        keep = dict(self.__dict__)
        keep["_pending_state"] = self.state_dict()["fused"] if self._flat is not None else None
        keep["_l2_dense"] = [self._l2_of_param.get(id(p), 0.0) for _, p in self.dense_named]
        for k in ("_flat", "_hist", "_l2_of_param", "reg_accum", "reg_accum_shard", "_optimizer_step_pre_hooks",
                  "_optimizer_step_post_hooks"):
            keep.pop(k, None)
        return keep

    def __setstate__(self, state):
        l2_dense = state.pop("_l2_dense")
        pending = state.pop("_pending_state")
        super().__setstate__(state)
        self._l2_of_param = {id(p): l2 for (_, p), l2 in zip(self.dense_named, l2_dense) if l2 != 0.0}
        self._flat = self._hist = self.reg_accum = self.reg_accum_shard = None
        self._hist_base, self._dirty = 0, False
        for ts in self.table_sets:
            ts.s1 = ts.s2 = ts.bitmap = ts.last = None
        if pending is not None and "opt_dev" in pending:
            self._pending_state = pending

    # -------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        if self._flat is not None and self._flat_valid():
            self._flat["g"].zero_()
        else:
            for _, p in self.dense_named:
                p.grad = None
        for ts in self.table_sets:
            ts.plan.stash = None
            for p in ts.params:
                p.grad = None

    @torch.no_grad()
    def step(self, closure=None, apply_l2=False, grad_scale=1.0):
        """One optimizer step.  apply_l2=True adds the regulariser gradient 2*l2*w inside the kernels (fused fit path);
        with apply_l2=False the gradients are used as they are (the caller back-propagated the reg loss itself)."""
        if closure is not None:
            raise NotImplementedError("closure is not supported")
        self.step_local()
        self.step_exchange()
        self.step_apply(apply_l2, grad_scale)
        self.step_barrier()
        self.maybe_flush()
        return None

    def maybe_flush(self):
        """Bound the replay backlog of the lazy table semantics (call between steps, never inside a CUDA-graph capture)."""
        if self._dirty and self.flush_interval and self.steps - self._last_flush_step >= self.flush_interval:
            self.flush()

    # The four phases of a step.  Single GPU: only step_apply does anything.  Multi-GPU: step_local and step_apply are pure kernel
    # sequences (capturable into CUDA graphs); step_exchange / step_barrier are the two NCCL collectives in between.
    def step_local(self):
        """Batch side of the sharded backward: this rank's (keys, row sums, owner ranges) into the exported exchange buffers."""
        self.prepare()
        if self.dist_ctx is not None:
            self.dist_ctx.sharded.reduce_local()

    def step_exchange(self):
        """Data-parallel gradient SUM (losses are reduction='sum'); also the barrier after which all exchange buffers are complete."""
        if self.dist_ctx is not None:
            self.dist_ctx.all_reduce_sum(self._flat["g"])

    def step_barrier(self):
        """Peers may read this shard / overwrite their exchange buffers only after every owner has applied its updates."""
        if self.dist_ctx is not None:
            self.dist_ctx.barrier(self.dist_ctx.sharded.device)

    def step_apply(self, apply_l2=False, grad_scale=1.0):
        L = N.lib()
        f = self._flat
        st = N.stream_ptr()
        cfg0 = self._cfg(0.0)
        ctx = self.dist_ctx
        lazy = self._lazy_active()
        if ctx is not None:
            sh = ctx.sharded
            sh.pull_segments()      # unique local rows touched by ANY rank's batch + their summed gradients
            if lazy:
                # owner-side catch-up of those rows, BEFORE the step counter advances (replays up to the completed steps)
                sh.catch_up_pulled(self._cfg(self.l2_sharded[0]), self._cfg(self.l2_sharded[1]), f["opt_dev"], self._hist, self._hist_base,
                                   self.reg_accum_shard)
        if lazy:
            if self.steps + 1 - self._hist_base >= self._hist_cap:      # history ring full: settle every row, start a new window
                self.flush()
                self._hist_base = self.steps
                self._publish_lazy()
            N.check(L.xdfm_opt_tick_hist(N.ptr(f["opt_dev"]), cfg0, N.ptr(self._hist), self._hist_cap, self._hist_base, st))
        else:
            N.check(L.xdfm_opt_tick(N.ptr(f["opt_dev"]), cfg0, st))
        if f["n"] > 0:
            N.check(L.xdfm_flat_opt(cfg0, N.ptr(f["opt_dev"]), f["n"], N.ptr(f["w"]), N.ptr(f["g"]), N.ptr(f["s1"]), N.ptr(f["s2"]),
                                    N.ptr(f["l2vec"]) if (apply_l2 and f["any_l2"]) else None, float(grad_scale),
                                    N.ptr(self.reg_accum), st))
        for ts in self.table_sets:
            plan = ts.plan
            cfg = self._cfg(ts.l2 if apply_l2 else 0.0)
            dense_grads = [p.grad for p in ts.params]
            if plan.stash is not None:
                (uniq, seg_off, pos, nseg, n), gsum = plan.stash
                dense_pass = 0 if self.sparse_embedding_update else 1
                if self.kind == "sgd" and cfg.l2 == 0.0:
                    dense_pass = 0      # nothing moves on untouched rows: g == 0 and no optimizer state
                mark = False
                if dense_pass and lazy:
                    dense_pass, mark = 0, True      # untouched rows are replayed later (catch_up / flush)
                with ops.timed("rows_opt"):
                    N.check(L.xdfm_rows_opt(cfg, N.ptr(f["opt_dev"]), N.ptr_array([p.data for p in ts.params]),
                                            N.ptr_array(ts.s1) if ts.s1 else None, N.ptr_array(ts.s2) if ts.s2 else None,
                                            plan._c_row_off, plan.T, plan.width, N.ptr(uniq), N.ptr(gsum), N.ptr(nseg), n,
                                            float(grad_scale), N.ptr(ts.bitmap), dense_pass, N.ptr(self.reg_accum), st))
                    if mark:
                        N.check(L.xdfm_rows_mark_current(N.ptr(ts.last), N.ptr(uniq), N.ptr(nseg), n, N.ptr(f["opt_dev"]), st))
                        self._dirty = True
                plan.stash = None
            elif any(g is not None for g in dense_grads):
                # dense .grad tensors (generic autograd path): every table is one flat update
                for i, p in enumerate(ts.params):
                    if p.grad is None:
                        continue
                    l2v = None
                    N.check(L.xdfm_flat_opt(cfg, N.ptr(f["opt_dev"]), p.numel(), N.ptr(p.data), N.ptr(p.grad.contiguous()),
                                            N.ptr(ts.s1[i]) if ts.s1 else None, N.ptr(ts.s2[i]) if ts.s2 else None, l2v,
                                            float(grad_scale), None, st))
        if ctx is not None:
            sh = ctx.sharded
            dense_pass = 0 if (self.sparse_embedding_update or lazy) else 1
            sh.apply_optimizer(self._cfg(self.l2_sharded[0] if apply_l2 else 0.0), self._cfg(self.l2_sharded[1] if apply_l2 else 0.0),
                               f["opt_dev"], grad_scale, dense_pass, self.reg_accum_shard)
            if lazy:
                sh.mark_pulled(f["opt_dev"])
                self._dirty = True
                self._publish_lazy()
        self.steps += 1

    # ---- lazy dense-table semantics -------------------------------------------------------------------
    def _lazy_active(self):
        if not self.lazy_tables or self.sparse_embedding_update:
            return False
        if self.dist_ctx is not None:
            return not (self.kind == "sgd" and self.l2_sharded == (0.0, 0.0))
        if not self.table_sets:
            return False
        return not (self.kind == "sgd" and all(ts.l2 == 0.0 for ts in self.table_sets))

    def _publish_lazy(self):
        """Row-sharded tables: tell the lookup which history to replay stale rows with (None = rows are always current)."""
        if self.dist_ctx is None:
            return
        sh = self.dist_ctx.sharded
        if self._lazy_active() and self._flat is not None:
            sh.lazy = dict(cfg_emb=self._cfg(self.l2_sharded[0]), cfg_lin=self._cfg(self.l2_sharded[1]), opt_dev=self._flat["opt_dev"],
                           hist=self._hist, hist_base=self._hist_base)
        else:
            sh.lazy = None

    def catch_up(self, plan, cache, ids):
        """Training forward: bring the rows this batch looks up to the current step before they are read."""
        if not self._dirty or not self._lazy_active() or ids.shape[0] == 0:
            return
        for ts in self.table_sets:
            if ts.plan is plan and ts.last is not None:
                uniq, seg_off, pos, nseg, n = cache.get(plan, ids)
                with ops.timed("rows_opt"):
                    N.check(N.lib().xdfm_rows_catchup(self._cfg(ts.l2), N.ptr(self._flat["opt_dev"]), N.ptr(self._hist), self._hist_base,
                                                      N.ptr_array([p.data for p in ts.params]), N.ptr_array(ts.s1) if ts.s1 else None,
                                                      N.ptr_array(ts.s2) if ts.s2 else None, N.ptr(ts.last), plan._c_row_off, plan.T,
                                                      plan.width, N.ptr(uniq), N.ptr(nseg), n, N.ptr(self.reg_accum), N.stream_ptr()))

    def flush(self):
        """Replay every postponed row update (before anything reads whole tables: predict, state_dict, end of an epoch)."""
        self._last_flush_step = self.steps
        if not self._dirty or self._flat is None:
            return
        if self.dist_ctx is not None:
            # collective: every rank settles its shard at the same step; peers replay stale rows from (w, m, v, last) when they look
            # them up, so nobody may read a shard while its owner rewrites it -> barrier before anyone's next lookup
            self.dist_ctx.sharded.flush_rows(self._cfg(self.l2_sharded[0]), self._cfg(self.l2_sharded[1]), self._flat["opt_dev"], self._hist,
                                             self._hist_base, self.reg_accum_shard)
            self.dist_ctx.barrier(self.dist_ctx.sharded.device)
        for ts in self.table_sets:
            if ts.last is None:
                continue
            with ops.timed("rows_opt"):
                N.check(N.lib().xdfm_rows_flush(self._cfg(ts.l2), N.ptr(self._flat["opt_dev"]), N.ptr(self._hist), self._hist_base,
                                                N.ptr_array([p.data for p in ts.params]), N.ptr_array(ts.s1) if ts.s1 else None,
                                                N.ptr_array(ts.s2) if ts.s2 else None, N.ptr(ts.last), ts.plan._c_row_off, ts.plan.T,
                                                ts.plan.width, N.ptr(self.reg_accum), N.stream_ptr()))
        self._dirty = False

    def pop_reg_loss(self):
        """Sum of l2*w^2 accumulated since the last call (device float64 -> python float; one sync).  Multi-GPU: the dense
        part is identical on every rank (counted once), the table part is summed over the owners."""
        self.flush()
        if self.dist_ctx is not None:
            t = self.reg_accum_shard.clone()
            self.dist_ctx.all_reduce_sum(t)
            v = float(self.reg_accum.item()) + float(t.item())
            self.reg_accum_shard.zero_()
        else:
            v = float(self.reg_accum.item())
        self.reg_accum.zero_()
        return v
