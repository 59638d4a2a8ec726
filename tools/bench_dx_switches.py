import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/xdeepfm-pytorch_b200")
from deepctr import _native as Nv
DEV = "cuda:0"
r8 = lambda x: (x + 7) // 8 * 8
L = Nv.lib()
def run(B, m, D, H, Hp, dbg, reps=5):
    g = torch.Generator().manual_seed(0)
    R = B * D
    x0t = (torch.randn(R, r8(m), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    xkt = x0t if Hp == m else (torch.randn(R, r8(2 * Hp), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    dyt = (torch.randn(R, r8(H), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(H, Hp * m, generator=g) / (Hp * m) ** 0.5).to(DEV)
    wt = torch.empty(L.xdfm_cin_bwd_dx_tc_wt_elems(m, Hp, H, D), dtype=torch.bfloat16, device=DEV)
    HpQ = (Hp + 15) // 16 * 16
    dxk = torch.empty(R, HpQ, device=DEV); dx0 = torch.zeros(R, r8(m), device=DEV)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    L.xdfm_cin_dx_set_debug(dbg)
    ts = []
    for r in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        Nv.check(L.xdfm_cin_bwd_dx_tc(Nv.ptr(dyt), Nv.ptr(x0t), Nv.ptr(xkt), xkt.shape[1], Nv.ptr(W), Nv.ptr(wt), B, m, Hp, H, D, Nv.ptr(dxk), Nv.ptr(dx0), Nv.stream_ptr()))
        e1.record(); torch.cuda.synchronize()
        if r >= 2: ts.append(e0.elapsed_time(e1))
    L.xdfm_cin_dx_set_debug(0)
    print("dX Hp=%d debug=%d: %.3f ms" % (Hp, dbg, sorted(ts)[len(ts)//2]), flush=True)
if os.environ.get("DX_ONLY"):
    run(8192, 26, 16, 200, int(os.environ["DX_ONLY"]), 0)
    sys.exit(0)
for Hp in (100, 26):
    for dbg in (0, 1, 2, 3, 4, 7):
        run(8192, 26, 16, 200, Hp, dbg)
print("--- cluster / pair variants (debug 0)")
for cl in (1, 2):
    L.xdfm_cin_tc_set_cluster(cl)
    for pair in (0, 1):
        L.xdfm_cin_dx_set_pair(pair)
        print("cluster=%d pair=%d" % (cl, pair), end=" ")
        run(8192, 26, 16, 200, 100, 0)
        print("cluster=%d pair=%d" % (cl, pair), end=" ")
        run(8192, 22, 32, 256, 128, 0)
L.xdfm_cin_dx_set_pair(0); L.xdfm_cin_tc_set_cluster(2)
