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
        print("dX Hp=%d debug=%2d cluster=%d ns<=%s: %.3f ms" % (Hp, dbg, cl, os.environ.get("XDFM_DEBUG_DX_NS", "-"), sorted(ts)[len(ts) // 2]),
              flush=True)

    if os.environ.get("DX_PACE"):
        for dbg in (1024, 1024 + 2048, 1024 + 4096):
            run_dx(8192, 26, 16, 200, 100, dbg, 2)
            run_dx(8192, 26, 16, 200, 26, dbg, 2)
        return
    if os.environ.get("DX_SHAPES"):
        # cfg2 / cfg4 / cfg5-like layer shapes
        for (B, m, D, H, Hp) in ((8192, 26, 16, 200, 100), (8192, 26, 16, 200, 26), (8192, 22, 32, 256, 128), (8192, 22, 32, 256, 22),
                                 (4096, 26, 64, 200, 100), (4096, 26, 64, 200, 26), (8192, 39, 16, 128, 64), (8192, 10, 16, 96, 48)):
            print("B=%d m=%d D=%d H=%d" % (B, m, D, H), end=" ")
            run_dx(B, m, D, H, Hp, 0, int(os.environ.get("DX_CLUSTER", "2")))
        return
    if os.environ.get("DX_TRACE"):
        # clock stamps of CTA 0's hand-offs: where does a field's round trip go?
        tr = torch.zeros(2 * 32 * 16, dtype=torch.int64, device=DEV)
        for Hp, dbg in ((100, 1024), (100, 1024 + 3), (100, 1024 + 3 + 8), (100, 1024 + 8), (100, 1024 + 16 + 64)):
            tr.zero_()
            L.xdfm_cin_dx_set_trace(Nv.ptr(tr))
            run_dx(8192, 26, 16, 200, Hp, dbg, int(os.environ.get("DX_CLUSTER", "2")), reps=1)
            L.xdfm_cin_dx_set_trace(None)
            t = tr.cpu().view(2, 32, 16)
            t0 = int(t[0, 0, 0])
            print("Hp=%d debug=%d: rows = tile.group; columns = mma:acc_empty mma:w_full mma:committed tma:w_empty row4:acc_full | "
                  "arrive of row warps 4..11 (cycles since the first stamp)" % (Hp, dbg))
            for it in range(2):
                for g in range(26):
                    if int(t[it, g, 2]) == 0 or (it == 1 and g > 1) or (it == 0 and (g < 16 or g > 22)):
                        continue
                    print("%d.%02d " % (it, g) + " ".join("%7d" % ((int(v) - t0) % (1 << 32) if int(v) else -1) for v in t[it, g]))
        return
    if os.environ.get("DX_SKEL"):
        # which part of the skeleton of the version-2 kernel costs what (wide and narrow cfg2 layers)
        for Hp in (100, 26):
            for dbg in (0, 7, 7 + 8, 7 + 16, 7 + 32, 7 + 64, 7 + 8 + 16 + 32 + 64, 8, 16, 32, 64):
                run_dx(8192, 26, 16, 200, Hp, dbg, int(os.environ.get("DX_CLUSTER", "2")))
        for m in (6, 13, 26, 52):
            for dbg in (0, 7, 127):
                print("m=%d" % m, end=" ")
                run_dx(8192, m, 16, 200, 100, dbg, 2)
        for B in (1184, 8192, 16384):
            for dbg in (0, 7, 127):
                print("B=%d" % B, end=" ")
                run_dx(B, 26, 16, 200, 100, dbg, 2)
        return
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
        for ns in (("",) if os.environ.get("DX_SHAPES") or os.environ.get("DX_PACE") else ("",) if os.environ.get("DX_TRACE") else ("", "2") if os.environ.get("DX_SKEL") else ("", "2", "3")):
            env = dict(os.environ)
            if ns:
                env["XDFM_DEBUG_DX_NS"] = ns
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
