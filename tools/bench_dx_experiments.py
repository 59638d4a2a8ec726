"""Switch-off experiments on the dX kernel (round 2): which part of the per-field pipeline sets the pace?
debug bits: 1 no contraction FMAs, 2 no TMEM loads, 4 no MMAs, 8 no W'' stream (barrier hand-offs only), 16 no tile outputs.
XDFM_DEBUG_DX_NS=<n> (env, per process) caps the W'' ring depth."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "xdeepfm-pytorch_b200"))
    import torch
    from deepctr import _native as Nv
    L = Nv.lib()
    DEV = "cuda:0"
    r8 = lambda x: (x + 7) // 8 * 8

    def run_dx(B, m, D, H, Hp, dbg, cl, reps=5):
        g = torch.Generator().manual_seed(0)
        R = B * D
        x0t = (torch.randn(R, r8(m), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
        xkt = x0t if Hp == m else (torch.randn(R, r8(2 * Hp), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
        dyt = (torch.randn(R, r8(H), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
        W = (torch.randn(H, Hp * m, generator=g) / (Hp * m) ** 0.5).to(DEV)
        wt = torch.empty(L.xdfm_cin_bwd_dx_tc_wt_elems(m, Hp, H, D), dtype=torch.bfloat16, device=DEV)
        HpQ = (Hp + 15) // 16 * 16
        dxk = torch.empty(R, HpQ, device=DEV)
        dx0 = torch.zeros(R, r8(m), device=DEV)
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
        print("dX Hp=%d debug=%2d cluster=%d ns<=%s: %.3f ms" % (Hp, dbg, cl, os.environ.get("XDFM_DEBUG_DX_NS", "-"), sorted(ts)[len(ts) // 2]),
              flush=True)

    if os.environ.get("DX_FIELDS"):
        # per-tile fixed cost vs per-field cost: same rows, different field counts
        for m in (6, 13, 26, 52):
            for dbg in (0, 31):
                print("m=%d" % m, end=" ")
                run_dx(8192, m, 16, 200, 100, dbg, 1)
        return
    if os.environ.get("DX_TILES"):
        # per-kernel fixed cost vs per-tile cost: same fields, 7 / 14 / 28 tiles per CTA
        for B in (1184, 8192, 16384, 32768):
            for m in (26, 13):
                for dbg in (0, 31):
                    print("B=%d m=%d" % (B, m), end=" ")
                    run_dx(B, m, 16, 200, 100, dbg, 1)
        return
    for Hp in (100, 26):
        for dbg in (0, 7, 7 + 8, 7 + 8 + 16, 8, 16):
            run_dx(8192, 26, 16, 200, Hp, dbg, 1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for ns in ("", "2", "3"):
            env = dict(os.environ)
            if ns:
                env["XDFM_DEBUG_DX_NS"] = ns
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
