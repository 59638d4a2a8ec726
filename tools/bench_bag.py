"""Micro-benchmark of the bag-pooling kernels (csrc/bag.cu) against the HBM roofline: CUDA events, working set >> L2 (126 MB).

usage (GPU box): python tools/bench_bag.py > gpurun_out/bench_bag.log
Layout: 24 fixed fields + two 20-position sequence features ('mean' under the id mask, 'sum' under a length column) = 64 slots ->
26 fields, D = 16 and 64, B = 65536.  Algorithmic bytes per launch: forward read [B, S, D] + write [B, F, D] (+ ids [B, S] int32);
backward read [B, F, D] + write [B, S, D] (+ ids)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xdeepfm-pytorch_b200"))
from deepctr import ops  # noqa: E402

DEV = "cuda:0"


def timeit(fn, reps=7):
    ts = []
    for r in range(reps + 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if r >= 3:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    peak = 6539.5
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    fields = [(1, "single", -1)] * 24 + [(20, "mean", -1), (20, "sum", 0)]
    lay = ops.BagLayout(fields)
    B = 65536
    g = torch.Generator().manual_seed(0)
    ids = (torch.randint(1, 1000, (B, lay.S), generator=g) * (torch.rand(B, lay.S, generator=g) < 0.7)).to(torch.int32).to(DEV)
    lens = torch.randint(0, 21, (B, 1), generator=g).to(torch.int32).to(DEV)
    for D in (16, 64):
        emb = torch.randn(B, lay.S, D, device=DEV, requires_grad=True)
        out = ops.BagPool.apply(lay, emb, ids, lens)
        dout = torch.randn_like(out)
        fwd_bytes = (B * lay.S * D + B * lay.F * D) * 4 + B * lay.S * 4
        ms_f = timeit(lambda: ops.BagPool.apply(lay, emb.detach(), ids, lens))
        ms_b = timeit(lambda: torch.autograd.grad(out, emb, dout, retain_graph=True))
        print("bag_pool D=%d B=%d S=%d F=%d: fwd %.3f ms = %.0f GB/s (%.0f %% of %.0f), bwd %.3f ms = %.0f GB/s (%.0f %%); %.0f MB / launch"
              % (D, B, lay.S, lay.F, ms_f, fwd_bytes / ms_f / 1e6, 100 * fwd_bytes / ms_f / 1e6 / peak, peak, ms_b,
                 fwd_bytes / ms_b / 1e6, 100 * fwd_bytes / ms_b / 1e6 / peak, fwd_bytes / 1e6), flush=True)


if __name__ == "__main__":
    main()
