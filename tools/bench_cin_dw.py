"""Micro-benchmark of the tensor-core CIN weight-gradient kernel (xdfm_cin_bwd_dw_tc) at BASELINE config shapes: lane packing on / off.
CUDA events, L2 flushed between launches.   python tools/bench_cin_dw.py"""
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


def r16(x):
    return (x + 15) // 16 * 16


def time_dw(B, m, D, H, Hp, pack, reps=7):
    L = Nv.lib()
    L.xdfm_cin_dw_set_pack(pack)
    g = torch.Generator().manual_seed(0)
    R = B * D
    x0T = (torch.randn(r8(m), R, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    xkT = x0T if (Hp == m and r16(Hp) == r8(m)) else (torch.randn(r16(Hp), R, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    dyT = (torch.randn(r16(H), R, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    dW = torch.empty(H, Hp * m, device=DEV)
    db = torch.empty(H, device=DEV)
    nb = L.xdfm_cin_bwd_dw_tc_workspace_bytes(B, m, Hp, H, D)
    ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    ts = []
    for r in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        Nv.check(L.xdfm_cin_bwd_dw_tc(Nv.ptr(dyT), Nv.ptr(xkT), Nv.ptr(x0T), B, m, Hp, H, D, Nv.ptr(dW), None, Nv.ptr(ws), nb, Nv.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        if r >= 2:
            ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print("dW pack=%d B=%d m=%d D=%d H=%d Hp=%d: %.3f ms (kernel + split reduce)  %.1f TFLOP/s algorithmic" % (
        pack, B, m, D, H, Hp, ms, 2.0 * R * H * Hp * m / ms / 1e9), flush=True)
    L.xdfm_cin_dw_set_pack(1)


if __name__ == "__main__":
    for pack in (0, 1):
        time_dw(8192, 26, 16, 200, 26, pack)      # cfg2 layer 0
        time_dw(8192, 26, 16, 200, 100, pack)     # cfg2 layers 1, 2 (no packing possible)
        time_dw(8192, 22, 32, 256, 22, pack)      # cfg4 layer 0
        time_dw(16384, 26, 16, 256, 26, pack)     # cfg3 layer 0
