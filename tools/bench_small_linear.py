"""A/B micro-benchmark of the narrow-layer kernels (csrc/smalllin.cu) at BASELINE config-3 shape: R = 16384 * 256 rows, K = N = 16."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xdeepfm-pytorch_b200"))
from deepctr import _native as Nv  # noqa: E402

DEV = "cuda:0"
L = Nv.lib()
R, K, N = 16384 * 256, 16, 16
g = torch.Generator().manual_seed(0)
x = torch.randn(R, K, generator=g).to(DEV)
Ws = [torch.randn(N, K, generator=g).to(DEV) for _ in range(3)]
ys = [torch.empty(R, N, device=DEV) for _ in range(3)]
dx = torch.empty(R, K, device=DEV)
dWs = [torch.empty(N, K, device=DEV) for _ in range(3)]
ws = torch.empty(L.xdfm_small_linear_bwd_dw_workspace_bytes(R, K, N, 3), dtype=torch.uint8, device=DEV)
st = Nv.stream_ptr()


def timeit(fn, reps=5):
    ts = []
    for r in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if r >= 2:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


P = Nv.ptr
for staged in (0, 1, 2, 4):
    L.xdfm_small_linear_set_staged(staged)
    t1 = timeit(lambda: Nv.check(L.xdfm_small_linear_fwd(P(x), P(Ws[0]), None, None, None, 0, R, K, N, 1, P(ys[0]), None, None, st)))
    t3 = timeit(lambda: Nv.check(L.xdfm_small_linear_fwd(P(x), P(Ws[0]), P(Ws[1]), P(Ws[2]), None, 0, R, K, N, 3, P(ys[0]), P(ys[1]), P(ys[2]), st)))
    d1 = timeit(lambda: Nv.check(L.xdfm_small_linear_bwd_dx(P(ys[0]), None, None, P(Ws[0]), None, None, R, K, N, 1, P(dx), st)))
    d3 = timeit(lambda: Nv.check(L.xdfm_small_linear_bwd_dx(P(ys[0]), P(ys[1]), P(ys[2]), P(Ws[0]), P(Ws[1]), P(Ws[2]), R, K, N, 3, P(dx), st)))
    gb = R * K * 4 / 1e9
    print("staged=%d  fwd x1 %.3f ms (%.0f GB/s)  fwd x3 %.3f ms (%.0f GB/s)  dx x1 %.3f ms (%.0f GB/s)  dx x3 %.3f ms (%.0f GB/s)" % (
        staged, t1, 2 * gb / t1 * 1e3, t3, 4 * gb / t3 * 1e3, d1, 2 * gb / d1 * 1e3, d3, 4 * gb / d3 * 1e3), flush=True)
w1 = timeit(lambda: Nv.check(L.xdfm_small_linear_bwd_dw(P(x), P(ys[0]), None, None, R, K, N, 1, P(dWs[0]), None, None, None, P(ws), st)))
w3 = timeit(lambda: Nv.check(L.xdfm_small_linear_bwd_dw(P(x), P(ys[0]), P(ys[1]), P(ys[2]), R, K, N, 3, P(dWs[0]), P(dWs[1]), P(dWs[2]), None, P(ws), st)))
print("dW x1 %.3f ms (%.0f GB/s)  dW x3 %.3f ms (%.0f GB/s)" % (w1, 2 * gb / w1 * 1e3, w3, 4 * gb / w3 * 1e3))
