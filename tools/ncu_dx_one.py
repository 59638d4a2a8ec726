import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "xdeepfm-pytorch_b200"))
import torch
from deepctr import _native as Nv
L = Nv.lib(); DEV = "cuda:0"
r8 = lambda x: (x + 7) // 8 * 8
B, m, D, H = 8192, 26, 16, 200
for Hp in (100, 26):
    g = torch.Generator().manual_seed(0); R = B * D
    x0t = (torch.randn(R, r8(m), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    xkt = x0t if Hp == m else (torch.randn(R, r8(2 * Hp), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    dyt = (torch.randn(R, r8(H), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    W = (torch.randn(H, Hp * m, generator=g) / (Hp * m) ** 0.5).to(DEV)
    wt = torch.empty(L.xdfm_cin_bwd_dx_tc_wt_elems(m, Hp, H, D), dtype=torch.bfloat16, device=DEV)
    HpQ = (Hp + 15) // 16 * 16
    dxk = torch.empty(R, HpQ, device=DEV); dx0 = torch.zeros(2, R, r8(m), device=DEV)
    for _ in range(3):
        Nv.check(L.xdfm_cin_bwd_dx_tc(Nv.ptr(dyt), Nv.ptr(x0t), Nv.ptr(xkt), xkt.shape[1], Nv.ptr(W), Nv.ptr(wt), B, m, Hp, H, D, Nv.ptr(dxk), Nv.ptr(dx0), Nv.stream_ptr()))
    torch.cuda.synchronize()
