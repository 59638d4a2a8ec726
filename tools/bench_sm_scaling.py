"""Is the weight stream of the CIN forward / dX kernels bound per SM or chip-wide (L2)?  Same launch sized for 148, 74 and 37 SMs
(XDFM_DEBUG_SMS, one process each): constant time per TILE and CTA = per-SM bound; constant TOTAL time = chip-wide bound."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "xdeepfm-pytorch_b200"))
    import torch
    import bench_cin_tc as T
    from deepctr import _native as Nv
    L = Nv.lib()
    DEV = "cuda:0"
    r8 = lambda x: (x + 7) // 8 * 8
    for cl in (2, 1):
        T.time_layer(8192, 26, 16, 200, 100, 200, cl)

    def run_dx(B, m, D, H, Hp, dbg, cl, reps=5):
        g = torch.Generator().manual_seed(0)
        R = B * D
        x0t = (torch.randn(R, r8(m), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
        xkt = (torch.randn(R, r8(2 * Hp), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
        dyt = (torch.randn(R, r8(H), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
        W = (torch.randn(H, Hp * m, generator=g) / (Hp * m) ** 0.5).to(DEV)
        wt = torch.empty(L.xdfm_cin_bwd_dx_tc_wt_elems(m, Hp, H, D), dtype=torch.bfloat16, device=DEV)
        HpQ = (Hp + 15) // 16 * 16
        dxk = torch.empty(R, HpQ, device=DEV)
        dx0 = torch.zeros(2, R, r8(m), device=DEV)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
        L.xdfm_cin_dx_set_debug(dbg)
        L.xdfm_cin_tc_set_cluster(cl)
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
        L.xdfm_cin_dx_set_debug(0)
        L.xdfm_cin_tc_set_cluster(2)
        print("dX Hp=%d debug=%d cluster=%d: %.3f ms" % (Hp, dbg, cl, sorted(ts)[len(ts) // 2]), flush=True)

    for dbg in (0, 7):
        for cl in (2, 1):
            run_dx(8192, 26, 16, 200, 100, dbg, cl)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for sms in (148, 74, 37):
            print("==== XDFM_DEBUG_SMS=%d" % sms, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, XDFM_DEBUG_SMS=str(sms)))
