#!/usr/bin/env python
"""Benchmark of the xDeepFM training hot path (BASELINE.json metric: train samples/sec, Criteo-shape xDeepFM).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2]

One "step" = one full training step (fused gather -> CIN -> DNN -> head -> BCE -> backward -> sorted scatter-add ->
fused Adam incl. the reference's dense-table L2/Adam semantics) over one batch of synthetic Criteo-shaped input.
Prints ONE JSON line on rank 0 (see the task contract): `value` = device-resident throughput, `e2e` = the same step
through the public `train_on_batch` API with pinned-host inputs (H2D + D2H inside the timed region), `roofline` for the
dominant kernel (timed live with CUDA events; `frac_incl_layout` counts the layout kernels that only serve it), `hbm_kernels`
(embedding gather / sorted segmented scatter-add at the cfg5 shape against the measured HBM peak), `fit_e2e` / `predict_e2e`
(the reference's own entry points `model.fit` / `model.predict` on host arrays), `cpu_baseline` = the unmodified reference
(oracle/_ref, shipped by oracle/build_ref.py) on this box's host cores; at N > 1 `dp_parity` (N GPUs == 1 GPU on the same global
batches, checked BEFORE timing; non-zero exit on mismatch) and `extra_workloads` (N = 1: cfg3, cfg4; N > 1: cfg4 = BASELINE configs[3] data-parallel; cfg5 at N = 8).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "xdeepfm-pytorch_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# Kaggle Display-Advertising-Challenge cardinalities of C1..C26 (assumption stated in SURVEY.md 8d; the reference ships no data)
CRITEO_VOCAB = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992, 5461306, 10,
                5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "cfg2": dict(m=26, nd=13, D=16, cin=(200, 200, 200), dnn=(400, 400), batch=8192, vocab=CRITEO_VOCAB),
    # BASELINE.json configs[0] (reference's CPU-runnable case) -- parity-test shape, selectable for quick runs
    "cfg1": dict(m=26, nd=13, D=8, cin=(256, 128), dnn=(256, 256), batch=256, vocab=[min(v, 100000) for v in CRITEO_VOCAB]),
    # BASELINE.json configs[2]: xDeepFM + field self-attention over the CIN maps (xdftrain_attn.py), 4 heads, batch 16384
    "cfg3": dict(m=26, nd=13, D=16, cin=(256, 128), dnn=(256, 256), batch=16384, vocab=CRITEO_VOCAB, variant="attn", heads=4),
    # BASELINE.json configs[3]: xdeepfm_pro, Avazu-shape 22 sparse fields, emb_dim 32, CIN 256x4; SFG heads need capped vocabularies
    # (Linear(64, V_f) + full-vocab cross-entropy per field: SURVEY.md 7.7), here <= 1e4 rows/field
    "cfg4": dict(m=22, nd=0, D=32, cin=(256, 256, 256, 256), dnn=(256, 256), batch=8192,
                 vocab=[241, 8, 8, 4738, 7746, 27, 8553, 560, 37, 10000, 10000, 8252, 6, 5, 2627, 9, 10, 436, 5, 69, 173, 61],
                 variant="pro"),
    # BASELINE.json configs[4]: large-vocab xDeepFM, 26 tables up to 50 M rows x 64 dims (192.7 M rows = 50 GB of fp32 tables, 150 GB
    # with Adam moments): never materialised whole (deferred_tables), row-sharded over the GPUs of the run (needs --gpus >= 2),
    # global batch 65536 at 8 GPUs (8192 / GPU).  Optimizer semantics: only the rows a batch touches are updated
    # (sparse_embedding_update; the reference's every-row-every-step Adam + L2 cannot run at this size).
    "cfg5": dict(m=26, nd=13, D=64, cin=(256, 128), dnn=(256, 256), batch=8192, deferred=True, sparse_update=True,
                 vocab=[50000000, 40000000, 30000000, 20000000, 10000000, 10000000, 10000000, 5000000, 5000000, 5000000, 2000000,
                        2000000, 1000000, 1000000, 1000000, 500000, 100000, 50000, 10000, 5000, 1000, 500, 100, 50, 10, 4]),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sust=d.get("bf16_tflops_sustained", 1400.0), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class WorkloadSpec:
    """Shapes of one bench workload for the PRODUCT arm (no oracle/ or tests/ import on that arm)."""

    def __init__(self, w):
        self.sparse_names = ["C%d" % i for i in range(1, w["m"] + 1)]
        self.dense_names = ["I%d" % i for i in range(1, w["nd"] + 1)]
        self.vocab_sizes = list(w["vocab"])
        self.embedding_dim = w["D"]
        self.cin_layer_size = tuple(w["cin"])
        self.cin_split_half = True
        self.dnn_hidden_units = tuple(w["dnn"])
        self.variant = w.get("variant", "xdeepfm")
        self.num_heads = w.get("heads", 4)
        self.m, self.nd = w["m"], w["nd"]


def build_product_model(spec, device):
    """The product model through the reference's own constructor API (xdftrain.py:269-285, xdftrain_attn.py, xdftrain_pro.py)."""
    from deepctr import models as M
    from deepctr.inputs import DenseFeat, SparseFeat
    cols = [SparseFeat(n, v, spec.embedding_dim) for n, v in zip(spec.sparse_names, spec.vocab_sizes)] + \
           [DenseFeat(n, 1) for n in spec.dense_names]
    common = dict(dnn_hidden_units=spec.dnn_hidden_units, cin_layer_size=spec.cin_layer_size, cin_split_half=True,
                  cin_activation="relu", l2_reg_linear=1e-5, l2_reg_embedding=1e-5, l2_reg_dnn=0.0, l2_reg_cin=0.0, device=device)
    if spec.variant == "xdeepfm":
        return M.xDeepFM(cols, cols, **common)
    if spec.variant == "pro":
        from deepctr.xdeepfm_pro import xDeepFMPro
        return xDeepFMPro(cols, cols, use_sfg=True, sfg_weight=0.1, sfg_hidden_units=(128, 64), sfg_dropout=0.0,
                          sfg_positive_only=True, sfg_use_label_attention=True, use_autodis=False, **common)
    if spec.variant == "attn":
        return M.xDeepFMAttention(cols, cols, cin_num_heads=spec.num_heads, cin_use_layer_norm=True, cin_use_residual=True, **common)
    raise SystemExit("unknown variant %s" % spec.variant)


def make_spec(w, vocab_cap=None):
    """ModelSpec of the CPU legs (`--impl reference` / cpu_baseline): the only place bench.py touches oracle/."""
    from oracle.xdeepfm_oracle import ModelSpec
    vocab = [min(v, vocab_cap) if vocab_cap else v for v in w["vocab"]]
    return ModelSpec(sparse_names=["C%d" % i for i in range(1, w["m"] + 1)], vocab_sizes=vocab, embedding_dim=w["D"],
                     dense_names=["I%d" % i for i in range(1, w["nd"] + 1)], cin_layer_size=tuple(w["cin"]),
                     dnn_hidden_units=tuple(w["dnn"]), variant=w.get("variant", "xdeepfm"), num_heads=w.get("heads", 4))


def synth_batches(spec, batch, n_batches, seed):
    """Seeded synthetic Criteo-shaped batches: ids log-uniform (Zipf-like) per field, dense U[0,1), labels Bernoulli(0.25)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        cols = []
        for V in spec.vocab_sizes:
            u = torch.rand(batch, generator=g)
            cols.append(torch.clamp((float(V) ** u).long() - 1, 0, V - 1).to(torch.int32))
        ids = torch.stack(cols, 1).contiguous()
        dense = torch.rand(batch, spec.nd, generator=g)
        y = (torch.rand(batch, generator=g) < 0.25).float()
        out.append((ids, dense, y))
    return out


def cin_flops_per_sample(spec):
    """Algorithmic CIN FLOPs per sample, forward: 2*D*sum_k H_k*K_k; training = 3x."""
    m, D = spec.m, spec.embedding_dim
    prev, tot = m, 0
    for k, H in enumerate(spec.cin_layer_size):
        tot += H * prev * m
        prev = H // 2 if spec.cin_split_half else H
    return 2.0 * D * tot


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons DURING the timed region (NVML every 20 ms; nvidia-smi as a fallback)."""
    BITS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.sm_max, self.how = index, [], set(), False, None, "nvml"
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h, self.how = None, "nvidia-smi"

    def _sample_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                           capture_output=True, text=True, timeout=5)
        f = [x.strip() for x in r.stdout.strip().split(",")]
        self.samples.append(float(f[0]))
        self.sm_max = float(f[1])
        for nme, v in zip(names, f[2:]):
            if v.lower().startswith("active"):
                self.reasons.add(nme)

    def run(self):
        while not self.stop_flag:
            try:
                if self._h is not None:
                    nv = self._nv
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for nme, bit in self.BITS.items():
                        if mask & bit:
                            self.reasons.add(nme)
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.02 if self._h is not None else 0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "how": self.how}


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_steps(spec, batch, steps, warmup, seed=2025):
    """Time `steps` training steps (forward, BCE-sum + L2, backward, dense Adam) on the CPU.
    Uses the unmodified reference when /root/reference exists (kind 'reference'), else the oracle port ('port')."""
    from oracle import xdeepfm_oracle as O
    from oracle.ref_loader import reference_available
    torch.set_num_threads(os.cpu_count() or 1)
    batches = synth_batches(spec, batch, 2, seed)
    kind = "port"
    if reference_available():
        try:
            from oracle.ref_loader import load_reference
            load_reference_ok = True
            # the product package is also called `deepctr`; the reference arm runs in its own process (see main)
            load_reference()
            from oracle.make_golden import build_reference_model
            model = build_reference_model(spec)
            model.compile("adam", "binary_crossentropy")
            kind = "reference"
        except Exception as e:  # pragma: no cover
            print("reference import failed (%s); using the oracle port" % e, file=sys.stderr)
            kind = "port"
    if kind == "port":
        params = {k: v.requires_grad_(True) for k, v in O.make_params(spec, seed=1).items()}
        opt = torch.optim.Adam(list(params.values()))

    def one_step(i):
        ids, dense, y = batches[i % len(batches)]
        X = torch.cat([ids.float(), dense], 1)
        if kind == "reference":
            y_pred = model(X).squeeze()
            model.optim.zero_grad()
            loss = torch.nn.functional.binary_cross_entropy(y_pred, y, reduction="sum")
            total = loss + model.get_regularization_loss() + model.aux_loss
            total.backward()
            model.optim.step()
        else:
            opt.zero_grad()
            loss, total = O.train_loss(params, spec, X, y)
            total.backward()
            opt.step()
        return float(loss.detach())

    for i in range(warmup):
        one_step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        one_step(warmup + i)
    dt = time.perf_counter() - t0
    return dict(samples_per_s=batch * steps / dt, ms_per_step=1e3 * dt / steps, kind=kind, cores=torch.get_num_threads())


# ------------------------------------------------------------------------------------------------
def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = make_spec(w, args.ref_vocab_cap)
    sample_batch = args.ref_batch
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    r = cpu_reference_steps(spec, sample_batch, steps, warmup)
    sample = "%d-sample batches (of the %d-sample workload batch), vocab capped at %s rows/field, Adam, %d warm-up + %d timed steps" % (
        sample_batch, w["batch"], args.ref_vocab_cap, warmup, steps)
    cfg = workload_config(args, w)
    # what THIS arm ran (the bounded CPU sample of the workload), next to the workload it samples
    cfg["reference_arm"] = {"batch": sample_batch, "vocab_cap_rows_per_field": args.ref_vocab_cap,
                            "total_rows": int(sum(spec.vocab_sizes)), "dtype": "f32", "device": "cpu", "threads": r["cores"],
                            "implementation": "unmodified reference (oracle/_ref or /root/reference)" if r["kind"] == "reference"
                            else "oracle port (reference tree not found)",
                            "processes": 1}
    line = {"impl": "reference", "metric": "train samples/sec (Criteo-shape xDeepFM)", "value": r["samples_per_s"],
            "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
            "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, w):
    return {"workload": "%s: %s, %d sparse + %d dense, emb_dim %d, CIN %s, DNN %s, batch %d/GPU, Adam, "
                        "synthetic cardinalities (%.1fM rows), reference dense-table L2+Adam semantics" % (
                            args.workload, {"xdeepfm": "xDeepFM Criteo-shape", "attn": "xDeepFMAttention (4-head self-attention over the CIN "
                                            "maps) Criteo-shape", "pro": "xDeepFMPro (SFG decoder) Avazu-shape"}[w.get("variant", "xdeepfm")],
                            w["m"], w["nd"], w["D"], tuple(w["cin"]), tuple(w["dnn"]), w["batch"], sum(w["vocab"]) / 1e6),
            "batch_per_gpu": w["batch"], "global_batch": w["batch"] * int(os.environ.get("WORLD_SIZE", "1")),
            "parallelism": "1 GPU" if int(os.environ.get("WORLD_SIZE", "1")) == 1 else
            "dp%d dense (NCCL all-reduce) + tables row-sharded x%d over NVLink peer memory" % (
                int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("WORLD_SIZE", "1"))),
            "optimizer": "adam", "cin_precision": args.cin_impl,
            "table_semantics": "only rows touched by a batch are updated (tables never materialised whole; NOT the reference's dense "
                               "semantics, which cannot run at this vocabulary size)" if w.get("sparse_update") else
                               "reference dense L2+Adam over every row, evaluated lazily (bit-identical replay of untouched rows; all "
                               "postponed updates are flushed inside the timed region)" if not args.dense_table_pass else
                               "reference dense L2+Adam over every row, streamed every step",
            "l2_flush": "not needed: 4 rotating batches; each step touches ~1 GB of fresh activations / gradients / random table rows "
                        "(>> 126 MB L2) and the flush streams all tables + Adam state (%.1f GB)" % (
                            sum(w["vocab"]) * (w["D"] + 1) * 4 * 3 / 1e9)}


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
class Env:
    """Process-wide state of one bench run: rank / world / device, barrier + max-over-ranks helpers."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        self.dev = "cuda:%d" % self.local_rank
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device(self.dev))
            self.dist = dist

    def sync_all(self):
        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        if self.dist is None:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()


def set_precision(model, cin_impl):
    model.cin.precision = cin_impl
    if hasattr(model, "dnn"):
        model.dnn.precision = "bf16" if cin_impl == "bf16" else "fp32"      # tcgen05 dense layers in the bf16 configuration
    if getattr(model, "sfg_decoder", None) is not None:
        model.sfg_decoder.precision = "bf16" if cin_impl == "bf16" else "fp32"


def build_workload_model(env, args, w, spec):
    B = w["batch"]
    if w.get("deferred"):
        if env.world < 2:
            raise SystemExit("workload keeps its tables row-sharded over the GPUs of the run: launch with --gpus >= 2 (torchrun)")
        from deepctr.inputs import deferred_tables
        with deferred_tables():
            model = build_product_model(spec, env.dev)
    else:
        model = build_product_model(spec, env.dev)
    # reference initialisation (init_std=1e-4 embeddings, default-init CIN) is what a user trains from
    if env.world > 1:
        # hybrid parallel: tables row-sharded over NVLink peer memory, dense part data-parallel (deepctr/distributed.py)
        model.distribute(max_batch=B)
    model.compile("adam", "binary_crossentropy")
    set_precision(model, args.cin_impl)
    model.optim.lazy_tables = not args.dense_table_pass
    if w.get("sparse_update"):
        model.optim.sparse_embedding_update = True
    return model


def release_model(model):
    """Free a workload's device memory before the next one (row-sharded tables are cudaMalloc'ed peer buffers)."""
    ctx = getattr(model, "_dist", None)
    if ctx is not None:
        # every rank unmaps its peers' buffers BEFORE any owner frees one (cudaFree of a region another process still has open is
        # undefined behaviour): barrier, close the mappings, barrier, then drop the owners
        torch.cuda.synchronize()
        ctx.barrier(ctx.sharded.device)
        torch.cuda.synchronize()
        for buf in (ctx.sharded.tables, ctx.sharded.exchange):
            if buf is not None:
                buf.close()
        torch.cuda.synchronize()
        ctx.barrier(ctx.sharded.device)
        torch.cuda.synchronize()
    model._graphs.clear()
    del model
    import gc
    gc.collect()
    torch.cuda.empty_cache()


def run_workload(env, args, name, steps, want_profile, want_e2e=True):
    """Warm up, (optionally) time every operator on eager launches, then time `steps` graph-replayed steps device-resident and
    end to end.  Returns a dict of raw measurements (rank-local except the max-over-ranks times)."""
    from deepctr import _native, ops
    w = WORKLOADS[name]
    spec = WorkloadSpec(w)
    B = w["batch"]
    model = build_workload_model(env, args, w, spec)
    n_pool = 4
    host = [(i.pin_memory(), d.pin_memory(), y.pin_memory()) for i, d, y in synth_batches(spec, B, n_pool, seed=2025 + env.rank)]
    devb = [(i.to(env.dev), d.to(env.dev), y.to(env.dev)) for i, d, y in host]
    # xDeepFM Pro sizes its positive-rows-only SFG pass from the labels' host copy (a count, no device sync); other models ignore it
    hostl = [y for _, _, y in host] if w.get("variant") == "pro" else [None] * n_pool
    accum = torch.zeros(1, dtype=torch.float64, device=env.dev)
    model.train()
    out = {"workload": name, "B": B, "spec": spec, "w": w}
    # ---- warm-up, eager launches
    graph_wanted = model.use_cuda_graph
    model.use_cuda_graph = False
    for i in range(max(args.warmup, 3)):
        model.train_step(*devb[i % n_pool], accum, host_labels=hostl[i % n_pool])
    env.sync_all()
    if want_profile:
        # ---- per-kernel-group device times (CUDA events around every operator; eager launches, same step, same data).  The pass
        # runs for >= 2 s so that the clocks it sees are those of a long run; they are sampled and decide the roofline denominator
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(3):
            model.train_step(*devb[i % n_pool], accum, host_labels=hostl[i % n_pool])
        e1.record()
        env.sync_all()
        est = max(e0.elapsed_time(e1) / 3, 0.05)
        est = env.max_over_ranks([est])[0]      # every rank must run the SAME number of profiled steps (the steps hold collectives)
        prof_steps = int(min(2000, max(20, min(steps, 20), args.profile_seconds * 1e3 / est)))
        sampler = ClockSampler(env.local_rank)
        if env.rank == 0:
            sampler.start()
        ops.TIMERS = {}
        l0 = _native.lib().xdfm_launch_count()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for i in range(prof_steps):
            model.train_step(*devb[i % n_pool], accum, host_labels=hostl[i % n_pool])
        p1.record()
        env.sync_all()
        sampler.stop_flag = True
        out["prof_ms"], out["prof_steps"] = p0.elapsed_time(p1), prof_steps
        out["launches_per_step"] = (_native.lib().xdfm_launch_count() - l0) // prof_steps
        out["timers"] = ops.timer_totals()
        out["prof_clocks"] = sampler.summary()
        ops.TIMERS = None
    # ---- the step captures itself into a CUDA graph on the third call with the same shapes
    model.use_cuda_graph = graph_wanted
    for i in range(3 * n_pool):          # every distinct step shape (Pro: positive-row bucket) is seen three times -> captured
        model.train_step(*devb[i % n_pool], accum, host_labels=hostl[i % n_pool])
    model.optim.flush()
    env.sync_all()
    out["graphed"] = bool(model._graphs)
    # ---- timed: device-resident inputs
    sampler = ClockSampler(env.local_rank)
    if env.rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.sync_all()
    e0.record()
    t_host0 = time.perf_counter()
    for i in range(steps):
        model.train_step(*devb[i % n_pool], accum, host_labels=hostl[i % n_pool])
    model.optim.flush()      # lazy dense-table semantics: every postponed row update is replayed INSIDE the timed region
    out["host_enqueue_ms"] = 1e3 * (time.perf_counter() - t_host0) / steps      # Python + launch time per step (no sync inside)
    e1.record()
    env.sync_all()
    ms = e0.elapsed_time(e1)
    ms_e2e = float("nan")
    if want_e2e:
        # ---- timed: end-to-end through the public API with pinned host inputs
        for i in range(n_pool):
            model.train_on_batch(*host[i % n_pool])
        env.sync_all()
        t0 = time.perf_counter()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for i in range(steps):
            model.train_on_batch(*host[i % n_pool])
        model.optim.flush()
        e3.record()
        env.sync_all()
        ms_e2e = max(e2.elapsed_time(e3), 1e3 * (time.perf_counter() - t0))
    sampler.stop_flag = True
    out["clocks"] = sampler.summary()
    out["ms"], out["ms_e2e"] = env.max_over_ranks([ms, ms_e2e])
    out["steps"] = steps
    out["h2d"] = sum(t.numel() * t.element_size() for t in host[0])
    out["model"] = model
    return out


# ------------------------------------------------------------------------------------------------
# N GPUs == 1 GPU, checked on the hardware of the run before anything is timed
# ------------------------------------------------------------------------------------------------
def seeded_parameters(model, seed):
    """O(0.1..1) parameters (embeddings std 0.5, matrices 1/sqrt(fan_in)) so that every branch of the model matters for the check
    (the reference's init_std = 1e-4 makes the CIN / DNN contributions ~1e-8).  Identical on every rank."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        shp = tuple(v.shape)
        fan_in = shp[1] if len(shp) > 1 else max(shp[0], 1)
        if "embedding_dict" in k and not k.startswith("linear_model."):
            std = 0.5
        elif k.startswith("linear_model.") or k.endswith(".bias"):
            std = 0.1
        else:
            std = 1.0 / max(fan_in, 1) ** 0.5
        t = torch.randn(shp, generator=g) * std
        if "layer_norm" in k and k.endswith("weight"):
            t = 1.0 + 0.1 * t
        sd[k] = t.to(v.dtype)
    return sd


PARITY_WORKLOADS = {
    "xdeepfm": dict(m=8, nd=3, D=16, cin=(48, 32), dnn=(64, 48), vocab=[50, 7, 3000, 3, 29, 400, 1000, 12]),
    "pro": dict(m=6, nd=2, D=16, cin=(32, 32), dnn=(48, 32), vocab=[40, 9, 500, 3, 31, 200], variant="pro"),
}


def run_dp_parity(env, args, variant, steps=4, per_rank=256):
    """K steps of the hybrid-parallel model on N GPUs (each rank its slice of every global batch) and the same global batches on
    rank 0 alone; both in the benchmarked precision.  Semantics that must hold (reference: basemodel.py:206-209, 254: per-GPU batch,
    SUM not mean; basemodel_sfg.py:316-349 for xDeepFM Pro): per-step losses, every dense parameter, every table row.
    Dense weight gradients are sums over rows whose fp32 association differs between the splits, and Adam normalises: tolerance
    1e-2 of each tensor's movement (tables: same), losses 1e-5 relative."""
    from deepctr.distributed import rank_slice
    w = dict(PARITY_WORKLOADS[variant], batch=per_rank)
    spec = WorkloadSpec(w)
    gb = per_rank * env.world
    batches = synth_batches(spec, gb, steps, seed=4242)         # same on every rank
    model = build_product_model(spec, env.dev)
    params = seeded_parameters(model, 31)
    model.load_state_dict(params, strict=True)
    model.distribute(max_batch=per_rank)
    model.compile("adam", "binary_crossentropy")
    set_precision(model, args.cin_impl)
    model.train()
    accum = torch.zeros(1, dtype=torch.float64, device=env.dev)
    losses = []
    for ids, dense, y in batches:
        a, b = rank_slice(0, gb, env.rank, env.world)
        accum.zero_()
        model.train_step(ids[a:b].to(env.dev), dense[a:b].to(env.dev), y[a:b].to(env.dev), accum, host_labels=y[a:b])
        t = accum.clone()
        env.dist.all_reduce(t)
        losses.append(t.item())
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}          # collective: tables re-assembled from the shards
    release_model(model)
    res = None
    if env.rank == 0:
        ref = build_product_model(spec, env.dev)
        ref.load_state_dict(params, strict=True)
        ref.compile("adam", "binary_crossentropy")
        set_precision(ref, args.cin_impl)
        ref.train()
        ref_losses = []
        for ids, dense, y in batches:
            accum.zero_()
            ref.train_step(ids.to(env.dev), dense.to(env.dev), y.to(env.dev), accum, host_labels=y)
            ref_losses.append(accum.item())
        rsd = {k: v.detach().cpu() for k, v in ref.state_dict().items()}
        rel_loss = max(abs(a - b) / abs(b) for a, b in zip(losses, ref_losses))
        dense_rel, table_rel = 0.0, 0.0
        for k in rsd:
            moved = (rsd[k].double() - params[k].double()).norm().item()
            err = (sd[k].double() - rsd[k].double()).norm().item()
            rel = err / max(moved, 1e-12) if moved > 0 else (0.0 if err == 0 else float("inf"))
            if "embedding_dict" in k:
                table_rel = max(table_rel, rel)
            else:
                dense_rel = max(dense_rel, rel)
        res = {"model": variant, "world": env.world, "steps": steps, "global_batch": gb, "precision": args.cin_impl,
               "max_rel_loss": rel_loss, "max_rel_dense_param": dense_rel, "max_rel_table": table_rel,
               "tables_equal": bool(table_rel <= 1e-2), "losses": losses, "losses_1gpu": ref_losses,
               "tolerance": "losses 1e-5 relative; parameters / table rows 1e-2 of the tensor's movement over the steps (norm)",
               "ok": bool(rel_loss <= 1e-5 and dense_rel <= 1e-2 and table_rel <= 1e-2)}
        del ref
        torch.cuda.empty_cache()
    flag = torch.tensor([1.0 if (res is None or res["ok"]) else 0.0], device=env.dev)
    env.dist.all_reduce(flag, op=env.dist.ReduceOp.MIN)
    return res, bool(flag.item() > 0)


# ------------------------------------------------------------------------------------------------
# HBM-bound kernels at the cfg5 shape (north_star: achieved GB/s of gather / scatter against the measured HBM peak)
# ------------------------------------------------------------------------------------------------
def measure_hbm_kernels(env, peaks, B=65536, D=64):
    """Fused multi-table gather and the sorted segmented scatter-add at BASELINE configs[4]'s per-step shape (65 536 samples x 26
    fields x 64 dims; tables capped at 2 M rows so that they fit one GPU -- 26 tables, 3.0 GB, far beyond L2), through the C ABI,
    CUDA events, L2 flushed between launches (a 256 MB write)."""
    from deepctr import _native as Nv
    from deepctr import ops
    L = Nv.lib()
    dev = env.dev
    rows = [min(v, 2000000) for v in CRITEO_VOCAB]
    m = len(rows)
    tables = [torch.randn(v, D, device=dev) for v in rows]
    g = torch.Generator().manual_seed(0)
    ids = torch.stack([torch.clamp((float(v) ** torch.rand(B, generator=g)).long() - 1, 0, v - 1) for v in rows], 1).to(torch.int32).to(dev)
    plan = ops.SparsePlan(list(range(m)), rows, D)
    out = torch.empty(B, m, D, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, reps=7):
        ts = []
        for r in range(reps + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    def gather():
        Nv.check(L.xdfm_embed_gather(Nv.ptr_array(tables), None, plan._c_vocab, Nv.ptr(ids), B, m, D, Nv.ptr(out), None, 0, None, None,
                                     Nv.stream_ptr()))
    res = {}
    ms = timeit(gather)
    nbytes = B * m * (4 + 2 * D * 4)
    res["embed_gather"] = {"achieved": nbytes / ms / 1e6, "peak": peaks["hbm"], "unit": "GB/s", "frac": nbytes / ms / 1e6 / peaks["hbm"],
                           "ms": ms, "bytes": nbytes, "shape": "B=%d m=%d D=%d fp32, 26 tables <= 2M rows, log-uniform ids" % (B, m, D),
                           "bytes_are": "id (4 B) + table row read + output row write per looked-up row"}
    # backward: (a) keys -> radix sort -> run-length segments, (b) deterministic segmented reduce of the [B*m, D] gradient rows
    n = B * m
    demb = torch.randn(B, m, D, device=dev)
    uniq = torch.empty(n, dtype=torch.int32, device=dev)
    seg_off = torch.empty(n + 1, dtype=torch.int32, device=dev)
    pos = torch.empty(n, dtype=torch.int32, device=dev)
    nseg = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(int(L.xdfm_embed_bwd_workspace_bytes(n)), dtype=torch.uint8, device=dev)
    gsum = torch.empty(n, D, device=dev)

    def segments():
        Nv.check(L.xdfm_embed_bwd_segments(Nv.ptr(ids), B, m, plan._c_feat_off, plan._c_vocab, plan.row_off[-1], Nv.ptr(ws), ws.numel(),
                                           Nv.ptr(uniq), Nv.ptr(seg_off), Nv.ptr(pos), Nv.ptr(nseg), Nv.stream_ptr()))

    def reduce_():
        Nv.check(L.xdfm_embed_bwd_reduce(Nv.ptr(demb), None, Nv.ptr(pos), Nv.ptr(seg_off), Nv.ptr(nseg), n, m, D, Nv.ptr(gsum), None,
                                         Nv.stream_ptr()))
    ms_seg = timeit(segments)
    ms_red = timeit(reduce_)
    u = int(nseg.item())
    red_bytes = n * D * 4 + u * D * 4 + n * 4 + (u + 1) * 4
    res["embed_scatter_reduce"] = {"achieved": red_bytes / ms_red / 1e6, "peak": peaks["hbm"], "unit": "GB/s",
                                   "frac": red_bytes / ms_red / 1e6 / peaks["hbm"], "ms": ms_red, "bytes": red_bytes,
                                   "unique_rows": u, "bytes_are": "gradient rows read (n*D*4) + unique-row sums written (u*D*4) + "
                                   "sorted positions and segment offsets read"}
    tot_bytes = red_bytes + n * 4
    res["embed_scatter"] = {"achieved": tot_bytes / (ms_seg + ms_red) / 1e6, "peak": peaks["hbm"], "unit": "GB/s",
                            "frac": tot_bytes / (ms_seg + ms_red) / 1e6 / peaks["hbm"], "ms": ms_seg + ms_red, "ms_sort_segments": ms_seg,
                            "bytes": tot_bytes, "note": "whole backward of the lookup: make keys + cub radix sort (only the key bits in "
                            "use) + run-length encode + scan, then the segmented reduce; the sort moves keys and positions several "
                            "times, bytes counted are the algorithmic ones only"}
    del tables, out, demb, gsum, flush
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------
# the reference's own entry points: model.fit / model.predict on host arrays
# ------------------------------------------------------------------------------------------------
def measure_fit_predict(env, args, w, n_rows=2 ** 21):
    """samples/s through `model.fit(x, y, batch_size, epochs=1, shuffle=True, verbose=0)` and `model.predict(x, batch_size)` on
    `n_rows` synthetic rows given as a dict of numpy arrays (xdftrain.py:444-458); the first epoch / call warms up (graph capture,
    pinned staging), the second is timed by wall clock around the call (everything inside: host batching, H2D, D2H)."""
    spec = WorkloadSpec(w)
    model = build_product_model(spec, env.dev)
    model.compile("adam", "binary_crossentropy")
    set_precision(model, args.cin_impl)
    model.optim.lazy_tables = not args.dense_table_pass
    B = w["batch"]
    g = np.random.default_rng(7)
    x = {}
    for name, V in zip(spec.sparse_names, spec.vocab_sizes):
        x[name] = np.minimum((float(V) ** g.random(n_rows)).astype(np.int64) - 1, V - 1).clip(0).astype(np.int32)
    for name in spec.dense_names:
        x[name] = g.random(n_rows, dtype=np.float32)
    y = (g.random(n_rows) < 0.25).astype(np.float32).reshape(-1, 1)
    import contextlib
    import io
    res = {}
    with contextlib.redirect_stdout(io.StringIO()):
        model.fit(x, y, batch_size=B, epochs=1, verbose=0, shuffle=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hist = model.fit(x, y, batch_size=B, epochs=1, verbose=0, shuffle=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    res["fit_e2e"] = {"value": n_rows / dt, "unit": "samples/s", "rows": n_rows, "batch_size": B, "seconds": dt, "epoch_loss": hist.history["loss"][0],
                      "how": "wall clock around model.fit(dict of numpy arrays, y, batch_size, epochs=1, shuffle=True, verbose=0), second epoch"}
    n_pred = n_rows // 2
    xp = {k: v[:n_pred] for k, v in x.items()}
    with contextlib.redirect_stdout(io.StringIO()):
        model.predict(xp, batch_size=B)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pred = model.predict(xp, batch_size=B)
        dt = time.perf_counter() - t0
    res["predict_e2e"] = {"value": n_pred / dt, "unit": "samples/s", "rows": n_pred, "batch_size": B, "seconds": dt,
                          "returns": "float64 [N, 1] numpy (as the reference)", "finite": bool(np.isfinite(pred).all()),
                          "how": "wall clock around model.predict(dict of numpy arrays, batch_size), second call"}
    release_model(model)
    return res


def load_cin_traffic():
    """DRAM bytes of the CIN contraction launches of one step, from the committed ncu --set full summary (written by
    tools/ncu_cin_traffic.py from an .ncu-rep of this repo's bench.py): (bytes, provenance) or (None, reason)."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_cin_traffic.json")
    if not os.path.exists(p):
        return None, "profiles/r02_ncu_cin_traffic.json not committed"
    try:
        d = json.load(open(p))
        return float(d["dram_bytes_per_step"]), "dram__bytes_read.sum + dram__bytes_write.sum over the %d CIN contraction launches of one " \
            "cfg2 step, ncu --set full, %s (csrc hash %s, captured %s)" % (d["launches_per_step"], os.path.relpath(p, ROOT), d.get("csrc_hash"), d.get("when"))
    except Exception as e:      # pragma: no cover
        return None, "unreadable %s: %s" % (p, e)


def roofline_block(args, peaks, r):
    """CIN contraction roofline from the operator timers of the profiling pass."""
    spec, B = r["spec"], r["B"]
    timers, prof_steps = r["timers"], r["prof_steps"]
    cin_ms = sum(timers.get(k, (0.0, 0))[0] for k in ("cin_fwd", "cin_bwd"))
    cin_calls = sum(timers.get(k, (0.0, 0))[1] for k in ("cin_fwd", "cin_bwd"))
    layout_ms = timers.get("cin_layout", (0.0, 0))[0]
    flops_step = 3.0 * cin_flops_per_sample(spec) * B
    achieved = flops_step * prof_steps / (cin_ms / 1e3) / 1e12 if cin_ms > 0 else None
    achieved_l = flops_step * prof_steps / ((cin_ms + layout_ms) / 1e3) / 1e12 if cin_ms > 0 else None
    # denominator: the burst figure when the SM clock during the profiled pass sat at its maximum (the step does not draw the power
    # of a pure GEMM loop, so it is NOT clock-limited the way the sustained measurement is); the sustained figure otherwise
    clk = r.get("prof_clocks") or {}
    at_max = bool(clk.get("sm_mhz") and clk.get("sm_max_mhz") and clk["sm_mhz"] >= 0.97 * clk["sm_max_mhz"])
    peak = peaks["tf_burst"] if at_max else peaks["tf_sust"]
    traffic, traffic_note = (load_cin_traffic() if (r["workload"] == "cfg2" and args.cin_impl == "bf16") else (None, "no capture for this workload"))
    return {"bound": "tensor", "kernel": "CIN contraction (cin_fwd + cin_bwd launches, %d per step)" % (cin_calls // max(prof_steps, 1)),
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
            "frac_incl_layout": (achieved_l / peak) if achieved_l else None,
            "frac_vs_burst": (achieved / peaks["tf_burst"]) if achieved else None,
            "frac_vs_sustained": (achieved / peaks["tf_sust"]) if achieved else None,
            "frac_incl_layout_vs_burst": (achieved_l / peaks["tf_burst"]) if achieved_l else None,
            "layout_note": "frac_incl_layout adds the cin_layout group (row / channel-major copies, dY, dX0 re-layout: kernels that exist "
                           "only to feed the contraction) to the kernel time",
            "traffic": traffic, "traffic_note": traffic_note,
            "algorithmic_hbm_bytes": 0.62e9 if r["workload"] == "cfg2" else None,
            "peak_source": "%s MEASURED_PEAKS.json bf16 %s: median SM clock of the %.1f s profiled pass %s MHz of %s max" % (
                peaks["src"], "burst (clock at max)" if at_max else "sustained (clock below max)", r["prof_ms"] / 1e3,
                clk.get("sm_mhz"), clk.get("sm_max_mhz")),
            "profile_clocks": clk,
            "share_of_step": (cin_ms / prof_steps) / (r["ms"] / r["steps"]) if r["ms"] > 0 else None,
            "timed_how": "CUDA events around every operator over %d eager steps (%.2f s) before the timed region (%.3f ms/step incl. event "
                         "overhead); share_of_step = that per-step kernel time / the timed region's ms_per_step (CUDA graph replay: %s)" % (
                             prof_steps, r["prof_ms"] / 1e3, r["prof_ms"] / prof_steps, r["graphed"]),
            "other_ms_per_step": {k: v[0] / prof_steps for k, v in timers.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cin-impl", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--ref-batch", type=int, default=1024)
    ap.add_argument("--ref-vocab-cap", type=int, default=100000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip hbm_kernels / fit_e2e / predict_e2e / dp_parity / extra_workloads")
    ap.add_argument("--profile-seconds", type=float, default=2.0, help="length of the eager operator-timing pass")
    ap.add_argument("--dense-table-pass", action="store_true",
                    help="stream every table row every step instead of the (bit-identical) lazy replay of untouched rows")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, w)

    # rank 0's stdout carries exactly ONE JSON line: everything else written to file descriptor 1 during the run (NCCL prints its
    # version banner there when the first communicator is created) is sent to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    env = Env()
    rank, world = env.rank, env.world
    peaks = load_peaks()
    extras = not args.no_extras

    # ---- N GPUs == 1 GPU on the same global batches, before anything is timed
    dp_parity = None
    if world > 1 and extras:
        dp_parity = []
        for variant in ("xdeepfm", "pro"):
            res, ok = run_dp_parity(env, args, variant)
            if rank == 0:
                dp_parity.append(res)
                print("dp_parity %s: %s" % (variant, json.dumps(res)), file=sys.stderr)
            if not ok:
                if rank == 0:
                    json_out.write(json.dumps({"error": "dp_parity failed", "dp_parity": dp_parity}) + "\n")
                    json_out.flush()
                env.dist.destroy_process_group()
                sys.exit(3)

    r = run_workload(env, args, args.workload, args.steps, want_profile=True)
    spec, B = r["spec"], r["B"]
    release_model(r.pop("model"))

    extra_workloads = None
    if extras and args.workload == "cfg2":
        extra_workloads = {}
        # one GPU: the attention variant (cfg3) and xDeepFM Pro (cfg4 = BASELINE configs[3]); N > 1: cfg4 data-parallel, cfg5 at N = 8
        names = ["cfg3", "cfg4"] if world == 1 else ["cfg4"] + (["cfg5"] if world >= 8 else [])
        for name in names:
            x = run_workload(env, args, name, min(args.steps, 30), want_profile=False)
            release_model(x.pop("model"))
            if rank == 0:
                extra_workloads[name] = {
                    "metric": "train samples/sec", "value": x["B"] * world * x["steps"] / (x["ms"] / 1e3), "unit": "samples/s",
                    "ms_per_step": x["ms"] / x["steps"], "e2e_value": x["B"] * world * x["steps"] / (x["ms_e2e"] / 1e3),
                    "e2e_ms_per_step": x["ms_e2e"] / x["steps"], "steps": x["steps"], "graph_replay": x["graphed"],
                    "config": workload_config(argparse.Namespace(**dict(vars(args), workload=name)), WORKLOADS[name]), "clocks": x["clocks"]}
    if rank != 0:
        if env.dist is not None:
            env.dist.barrier()
        return
    ms, ms_e2e = r["ms"], r["ms_e2e"]
    value = B * world * args.steps / (ms / 1e3)
    e2e = B * world * args.steps / (ms_e2e / 1e3)
    roofline = roofline_block(args, peaks, r)
    timers = r["timers"]
    # secondary (HBM-bound) kernels: inside the step (tiny launches at this batch) and at the cfg5 shape (what the roofline is for)
    hbm = {}
    if "embed_gather" in timers and timers["embed_gather"][0] > 0:
        gb = B * spec.m * (4 + 2 * spec.embedding_dim * 4) * timers["embed_gather"][1]
        a = gb / (timers["embed_gather"][0] / 1e3) / 1e9
        hbm["embed_gather_in_step"] = {"achieved": a, "peak": peaks["hbm"], "unit": "GB/s", "frac": a / peaks["hbm"],
                                       "note": "id + row read + row write per looked-up row; %.0f MB per launch at this batch: latency-"
                                               "bound, see the cfg5-shape entries" % (gb / timers["embed_gather"][1] / 1e6)}
    if world == 1 and extras:
        try:
            hbm.update(measure_hbm_kernels(env, peaks))
        except Exception as e:      # pragma: no cover
            hbm["error"] = "cfg5-shape kernel measurement failed: %s" % e
    roofline["hbm_kernels"] = hbm
    launches_per_step = r["launches_per_step"]
    line = {"metric": "train samples/sec (Criteo-shape xDeepFM)", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.cin_impl == "fp32" else "bf16", "data": "synthetic",
            "config": workload_config(args, w), "clocks": r["clocks"],
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_note": "%d kernels of libxdfm_sm100a.so per step (counted on "
            "eager steps) x %d steps; in the timed region they run as nodes of a replayed CUDA graph: %s" % (
                launches_per_step, args.steps, r["graphed"]),
            "host_enqueue_ms_per_step": r["host_enqueue_ms"], "roofline": roofline}
    if dp_parity is not None:
        line["dp_parity"] = dp_parity
    if extra_workloads is not None:
        line["extra_workloads"] = extra_workloads
    if world == 1 and extras and args.workload in ("cfg2", "cfg1"):
        try:
            line.update(measure_fit_predict(env, args, w))
        except Exception as e:      # pragma: no cover
            line["fit_e2e"] = {"value": None, "error": str(e)}
    if not args.no_cpu_baseline and world == 1:
        # bounded CPU sample (~10 s of host work on the box's cores) in a separate process (the reference package shares the name
        # `deepctr` with the product)
        try:
            rr = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "40" if args.workload in ("cfg1", "cfg2") else "10", "--warmup", "3",
                                 "--workload", args.workload, "--ref-batch", str(args.ref_batch),
                                 "--ref-vocab-cap", str(args.ref_vocab_cap)], capture_output=True, text=True, timeout=600)
            ref = json.loads(rr.stdout.strip().splitlines()[-1])
            line["cpu_baseline"] = ref["cpu_baseline"]
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: %s" % e}
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if env.dist is not None:
        env.dist.barrier()


if __name__ == "__main__":
    main()
