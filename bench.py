#!/usr/bin/env python
"""Benchmark of the xDeepFM training hot path (BASELINE.json metric: train samples/sec, Criteo-shape xDeepFM).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2]

One "step" = one full training step (fused gather -> CIN -> DNN -> head -> BCE -> backward -> sorted scatter-add ->
fused Adam incl. the reference's dense-table L2/Adam semantics) over one batch of synthetic Criteo-shaped input.
Prints ONE JSON line on rank 0 (see the task contract): `value` = device-resident throughput, `e2e` = the same step
through the public `train_on_batch` API with pinned-host inputs (H2D + D2H inside the timed region), `roofline` for the
dominant kernel (timed live with CUDA events), `cpu_baseline` = the CPU oracle port on this box's host cores.
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
    r = cpu_reference_steps(spec, sample_batch, max(1, args.steps), max(1, min(args.warmup, 1)))
    sample = "%d-sample batches (of the %d-sample workload batch), vocab capped at %s rows/field, Adam, %d timed steps" % (
        sample_batch, w["batch"], args.ref_vocab_cap, max(1, args.steps))
    line = {"impl": "reference", "metric": "train samples/sec (Criteo-shape xDeepFM)", "value": r["samples_per_s"],
            "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, w),
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
    ap.add_argument("--dense-table-pass", action="store_true",
                    help="stream every table row every step instead of the (bit-identical) lazy replay of untouched rows")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, w)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # rank 0's stdout carries exactly ONE JSON line: everything else written to file descriptor 1 during the run (NCCL prints its
    # version banner there when the first communicator is created) is sent to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from deepctr import _native, ops
    peaks = load_peaks()
    spec = WorkloadSpec(w)
    B = w["batch"]
    if w.get("deferred"):
        if world < 2:
            raise SystemExit("workload %s keeps its tables row-sharded over the GPUs of the run: launch with --gpus >= 2 (torchrun)" % args.workload)
        from deepctr.inputs import deferred_tables
        with deferred_tables():
            model = build_product_model(spec, dev)
    else:
        model = build_product_model(spec, dev)
    # reference initialisation (init_std=1e-4 embeddings, default-init CIN) is what a user trains from
    if world > 1:
        # hybrid parallel: tables row-sharded over NVLink peer memory, dense part data-parallel (deepctr/distributed.py)
        model.distribute(max_batch=B)
    model.compile("adam", "binary_crossentropy")
    model.cin.precision = args.cin_impl
    if hasattr(model, "dnn"):
        model.dnn.precision = "bf16" if args.cin_impl == "bf16" else "fp32"      # tcgen05 dense layers in the bf16 configuration
    if getattr(model, "sfg_decoder", None) is not None:
        model.sfg_decoder.precision = "bf16" if args.cin_impl == "bf16" else "fp32"
    model.optim.lazy_tables = not args.dense_table_pass
    if w.get("sparse_update"):
        model.optim.sparse_embedding_update = True
    n_pool = 4
    host = [(i.pin_memory(), d.pin_memory(), y.pin_memory()) for i, d, y in synth_batches(spec, B, n_pool, seed=2025 + rank)]
    devb = [(i.to(dev), d.to(dev), y.to(dev)) for i, d, y in host]
    # xDeepFM Pro sizes its positive-rows-only SFG pass from the labels' host copy (a count, no device sync); other models ignore it
    hostl = [y for _, _, y in host] if w.get("variant") == "pro" else [None] * n_pool
    accum = torch.zeros(1, dtype=torch.float64, device=dev)
    model.train()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up, eager launches
    graph_wanted = model.use_cuda_graph
    model.use_cuda_graph = False
    for i in range(max(args.warmup, 3)):
        model.train_step(*devb[i % n_pool], accum, host_labels=hostl[i % n_pool])
    sync_all()
    # ---- per-kernel-group device times (CUDA events around every operator; eager launches, same step, same data)
    prof_steps = min(args.steps, 20)
    ops.TIMERS = {}
    l0 = _native.lib().xdfm_launch_count()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(prof_steps):
        model.train_step(*devb[i % n_pool], accum, host_labels=hostl[i % n_pool])
    p1.record()
    sync_all()
    prof_ms = p0.elapsed_time(p1)
    launches_per_step = (_native.lib().xdfm_launch_count() - l0) // prof_steps
    launches = launches_per_step * args.steps
    timers = ops.timer_totals()
    ops.TIMERS = None
    # ---- the step captures itself into a CUDA graph on the third call with the same shapes
    model.use_cuda_graph = graph_wanted
    for i in range(3 * n_pool):          # every distinct step shape (Pro: positive-row bucket) is seen three times -> captured
        model.train_step(*devb[i % n_pool], accum, host_labels=hostl[i % n_pool])
    model.optim.flush()
    sync_all()
    graphed = bool(model._graphs)
    # ---- timed: device-resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        model.train_step(*devb[i % n_pool], accum, host_labels=hostl[i % n_pool])
    model.optim.flush()      # lazy dense-table semantics: every postponed row update is replayed INSIDE the timed region
    host_enqueue_ms = 1e3 * (time.perf_counter() - t_host0) / args.steps      # Python + launch time per step (no sync inside)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    # ---- timed: end-to-end through the public API with pinned host inputs
    for i in range(n_pool):
        model.train_on_batch(*host[i % n_pool])
    sync_all()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        model.train_on_batch(*host[i % n_pool])
    model.optim.flush()
    e3.record()
    sync_all()
    ms_e2e = max(e2.elapsed_time(e3), 1e3 * (time.perf_counter() - t0))
    sampler.stop_flag = True
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank != 0:
        return
    value = B * world * args.steps / (ms / 1e3)
    e2e = B * world * args.steps / (ms_e2e / 1e3)
    # ---- roofline of the dominant kernel group: the CIN contraction (fwd + bwd)
    cin_ms = sum(timers.get(k, (0.0, 0))[0] for k in ("cin_fwd", "cin_bwd"))
    cin_calls = sum(timers.get(k, (0.0, 0))[1] for k in ("cin_fwd", "cin_bwd"))
    flops_step = 3.0 * cin_flops_per_sample(spec) * B
    achieved = flops_step * prof_steps / (cin_ms / 1e3) / 1e12 if cin_ms > 0 else None
    peak = peaks["tf_sust"]
    roofline = {"bound": "tensor", "kernel": "CIN contraction (cin_fwd + cin_bwd launches, %d per step)" % (cin_calls // max(prof_steps, 1)),
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                "traffic": 0.80e9 if args.workload == "cfg2" and args.cin_impl == "bf16" else None,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the 9 CIN contraction launches of one step, bytes, from "
                                "profiles/r01c_ncu_full_cfg2.md (ncu --set full); algorithmic HBM bytes of the group ~0.62e9",
                "peak_source": "%s MEASURED_PEAKS.json bf16 sustained (kernel timed inside a long step)" % peaks["src"],
                "share_of_step": (cin_ms / prof_steps) / (ms / args.steps) if ms > 0 else None,
                "timed_how": "CUDA events around every operator over %d eager steps before the timed region (%.3f ms/step incl. event "
                             "overhead); share_of_step = that per-step kernel time / the timed region's ms_per_step (CUDA graph "
                             "replay: %s)" % (prof_steps, prof_ms / prof_steps, graphed),
                "other_ms_per_step": {k: v[0] / prof_steps for k, v in timers.items()}}
    # secondary (HBM-bound) kernels, timed live in the same run: algorithmic bytes / CUDA-event time
    hbm = {}
    if "embed_gather" in timers and timers["embed_gather"][0] > 0:
        gb = B * spec.m * (4 + 2 * spec.embedding_dim * 4) * timers["embed_gather"][1]
        a = gb / (timers["embed_gather"][0] / 1e3) / 1e9
        hbm["embed_gather"] = {"achieved": a, "peak": peaks["hbm"], "unit": "GB/s", "frac": a / peaks["hbm"],
                               "note": "id + row read + row write per looked-up row; 15 MB per launch at this batch (latency-bound: see "
                                       "profiles/ for the 450 MB/launch cfg5-shape measurement)"}
    roofline["hbm_kernels"] = hbm
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    line = {"metric": "train samples/sec (Criteo-shape xDeepFM)", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.cin_impl == "fp32" else "bf16", "data": "synthetic",
            "config": workload_config(args, w), "clocks": sampler.summary(),
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "gpu_launches_note": "%d kernels of libxdfm_sm100a.so per step (counted on eager steps) x %d "
            "steps; in the timed region they run as nodes of a replayed CUDA graph: %s" % (launches_per_step, args.steps, graphed),
            "host_enqueue_ms_per_step": host_enqueue_ms, "roofline": roofline}
    if not args.no_cpu_baseline and world == 1:
        # bounded CPU sample (~10 s of host work on the box's cores) in a separate process (the reference package shares the name
        # `deepctr` with the product)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "40" if args.workload in ("cfg1", "cfg2") else "10", "--warmup", "1",
                                "--workload", args.workload, "--ref-batch", str(args.ref_batch),
                                "--ref-vocab-cap", str(args.ref_vocab_cap)], capture_output=True, text=True, timeout=600)
            ref = json.loads(r.stdout.strip().splitlines()[-1])
            line["cpu_baseline"] = ref["cpu_baseline"]
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: %s" % e}
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()


if __name__ == "__main__":
    main()
