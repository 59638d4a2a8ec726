"""Micro-benchmarks of the HBM-bound / GEMM kernels at cfg2 shapes (CUDA events; working sets >> L2 or L2 flushed)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xdeepfm-pytorch_b200"))
from deepctr import _native as Nv, ops  # noqa: E402
from deepctr.optim import FusedOptimizer, TableSet  # noqa: E402

DEV = "cuda:0"
CRITEO_VOCAB = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992, 5461306, 10,
                5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]


def timeit(fn, reps=5, flush=None):
    ts = []
    for r in range(reps + 2):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if r >= 2:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def bench_rows_opt(D=16):
    L = Nv.lib()
    rows = CRITEO_VOCAB
    params = [torch.nn.Parameter(torch.randn(v, D, device=DEV) * 1e-2) for v in rows]
    plan = ops.SparsePlan(list(range(len(rows))), rows, D)
    ts = TableSet(plan, params, 1e-5)
    dense = [("w", torch.nn.Parameter(torch.zeros(8, device=DEV)))]
    opt = FusedOptimizer("adam", dense, [ts], {})
    opt.prepare()
    B, m = 8192, len(rows)
    g = torch.Generator().manual_seed(0)
    ids = torch.stack([torch.randint(0, v, (B,), generator=g) for v in rows], 1).to(torch.int32).to(DEV)
    cache = ops.SegmentCache()
    seg = cache.get(plan, ids)
    gsum = torch.randn(B * m, D, device=DEV) * 1e-3
    nelem = sum(rows) * D
    for ver in (1, 2):
        L.xdfm_set_rows_opt_dense_version(ver)

        def step():
            plan.stash = (seg, gsum)
            opt.step(apply_l2=True)
        ms = timeit(step)
        print("rows_opt dense v%d (Adam, %d rows x %d): %.3f ms  %.1f GB/s (24 B/element)" % (ver, sum(rows), D, ms, nelem * 24 / ms / 1e6),
              flush=True)
    L.xdfm_set_rows_opt_dense_version(2)


def bench_gemm():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for (M, N, K) in [(8192, 400, 429), (8192, 400, 400), (8192, 429, 400), (400, 429, 8192), (400, 400, 8192)]:
        A, Bm = torch.randn(M, K, device=DEV), torch.randn(N, K, device=DEV)
        Ab, Bb = ops.cvt_bf16(A), ops.cvt_bf16(Bm)
        bias = torch.zeros(N, device=DEV)
        ms = timeit(lambda: ops.gemm_tc(Ab, Bb, M, N, K, bias, 1), flush=flush)
        ms_c = timeit(lambda: ops.cvt_bf16(A), flush=flush)
        ms_t = timeit(lambda: ops.cvt_bf16(A, transpose=True), flush=flush)
        C = torch.empty(M, N, device=DEV)
        ms_s = timeit(lambda: ops.gemm(0, 1, M, N, K, A, K, Bm, K, C, N, bias=bias, act=1), flush=flush)
        print("gemm M=%d N=%d K=%d: tcgen05 %.3f ms (%.1f TFLOP/s)  cvt %.3f ms  cvt^T %.3f ms | fp32 sgemm %.3f ms" % (
            M, N, K, ms, 2.0 * M * N * K / ms / 1e9, ms_c, ms_t, ms_s), flush=True)


def bench_gather(D=64, B=65536):
    L = Nv.lib()
    rows = [min(v, 2000000) for v in CRITEO_VOCAB]
    tables = [torch.randn(v, D, device=DEV) for v in rows]
    g = torch.Generator().manual_seed(0)
    ids = torch.stack([torch.randint(0, v, (B,), generator=g) for v in rows], 1).to(torch.int32).to(DEV)
    plan = ops.SparsePlan(list(range(len(rows))), rows, D)
    m = len(rows)
    out = torch.empty(B, m, D, device=DEV)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

    def run():
        Nv.check(L.xdfm_embed_gather(Nv.ptr_array(tables), None, plan._c_vocab, Nv.ptr(ids), B, m, D, Nv.ptr(out), None, 0, None, None,
                                     Nv.stream_ptr()))
    ms = timeit(run, flush=flush)
    nbytes = B * m * (4 + 2 * D * 4)
    print("embed_gather B=%d m=%d D=%d: %.3f ms  %.1f GB/s algorithmic (id + row read + row write)" % (B, m, D, ms, nbytes / ms / 1e6),
          flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["rows_opt", "gemm", "gather"]
    if "gemm" in which:
        bench_gemm()
    if "gather" in which:
        bench_gather()
        bench_gather(D=16, B=65536)
    if "rows_opt" in which:
        bench_rows_opt()
