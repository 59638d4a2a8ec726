"""Micro-benchmark of the tensor-core CIN forward kernel at BASELINE config shapes (CUDA events, L2 flushed between launches)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xdeepfm-pytorch_b200"))
from deepctr import _native as Nv  # noqa: E402

DEV = "cuda:0"


def r8(x):
    return (x + 7) // 8 * 8


def time_layer(B, m, D, H, Hp, Hprev, cluster, reps=7):
    L = Nv.lib()
    L.xdfm_cin_tc_set_cluster(cluster)
    g = torch.Generator().manual_seed(0)
    x0t = (torch.randn(B * D, r8(m), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    xkt = x0t if Hp == m else (torch.randn(B * D, r8(Hprev), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    K = Hp * m
    W = (torch.randn(H, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.zeros(H, device=DEV)
    wprime = torch.empty(L.xdfm_cin_tc_wprime_elems(m, Hp, H, D), dtype=torch.bfloat16, device=DEV)
    yt = torch.empty(B * D, r8(H), dtype=torch.bfloat16, device=DEV)
    pooled = torch.empty(B, H, device=DEV)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    ts = []
    for r in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        Nv.check(L.xdfm_cin_fwd_tc(Nv.ptr(x0t), Nv.ptr(xkt), xkt.shape[1], Nv.ptr(W), Nv.ptr(b), Nv.ptr(wprime), B, m, Hp, H, D, 1,
                                   Nv.ptr(yt), H // 2, Nv.ptr(pooled), None, H, 0, Nv.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        if r >= 2:
            ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    flops = 2.0 * B * D * H * K
    print("cluster=%d B=%d m=%d D=%d H=%d Hp=%d: %.3f ms  %.1f TFLOP/s (algorithmic)" % (cluster, B, m, D, H, Hp, ms, flops / ms / 1e9),
          flush=True)
    return ms


if __name__ == "__main__":
    clusters = [int(c) for c in sys.argv[1:]] or [1, 2, 4]
    for pair in (0, 1):
      Nv.lib().xdfm_cin_tc_set_pair(pair)
      print("pair kernels:", pair, flush=True)
      for c in clusters:
        time_layer(8192, 26, 16, 200, 26, 26, c)
        time_layer(8192, 26, 16, 200, 100, 200, c)
        time_layer(16384, 26, 16, 128, 128, 256, c)
        time_layer(8192, 22, 32, 256, 128, 256, c)


def time_dx(B, m, D, H, Hp, cluster=2, reps=5):
    """dX kernel alone, with the diagnostic switches (which part of the per-field pipeline dominates?)."""
    L = Nv.lib()
    L.xdfm_cin_tc_set_cluster(cluster)
    g = torch.Generator().manual_seed(0)
    R = B * D
    x0t = (torch.randn(R, r8(m), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    xkt = x0t if Hp == m else (torch.randn(R, r8(2 * Hp), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    dyt = (torch.randn(R, r8(H), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(H, Hp * m, generator=g) / (Hp * m) ** 0.5).to(DEV)
    wt = torch.empty(L.xdfm_cin_bwd_dx_tc_wt_elems(m, Hp, H, D), dtype=torch.bfloat16, device=DEV)
    HpQ = (Hp + 15) // 16 * 16
    dxk = torch.empty(R, HpQ, device=DEV)
    dx0 = torch.zeros(2, R, r8(m), device=DEV)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for dbg in (0,):
        L.xdfm_cin_dx_set_debug(0)
        ts = []
        for r in range(reps + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            Nv.check(L.xdfm_cin_bwd_dx_tc(Nv.ptr(dyt), Nv.ptr(x0t), Nv.ptr(xkt), xkt.shape[1], Nv.ptr(W), Nv.ptr(wt), B, m, Hp, H, D,
                                          Nv.ptr(dxk), Nv.ptr(dx0), Nv.stream_ptr()))
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print("dX B=%d m=%d D=%d H=%d Hp=%d %s: %.3f ms  %.1f TFLOP/s" % (B, m, D, H, Hp, "", ms,
                                                                         2.0 * R * H * Hp * m / ms / 1e9), flush=True)


if __name__ == "__main__" and os.environ.get("BENCH_DX"):
    time_dx(8192, 26, 16, 200, 100)
    time_dx(8192, 26, 16, 200, 26)
