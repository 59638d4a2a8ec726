"""SM-cycle latencies of the tcgen05 / mbarrier hand-off primitives (csrc/tc_selftest.cu: tc_latency_probe_kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xdeepfm-pytorch_b200"))
import torch
from deepctr import _native as Nv

names = ["idle commit -> wait", "arrive -> wait (same thread)", "13 MMAs (N=112) + commit -> wait", "13 MMAs issue only",
         "try_wait on a completed phase", "tcgen05.ld x16 + wait::ld", "round trip by arrive", "round trip by commit",
         "round trip by two commits", "arrive vs 8 full warps", "commit vs 8 full warps", "2 commits vs 8 warps, warp 0 waits",
         "same + fences / wait::ld / STS (row-warp skeleton)", "13 MMAs + commit -> wait, 8 warps draining TMEM meanwhile",
         "13 MMAs + commit -> wait, 8 warps doing FMAs meanwhile"]
out = torch.zeros(16, dtype=torch.int64, device="cuda:0")
for setmax in (0, 1):
    for rep in range(3):
        Nv.check(Nv.lib().xdfm_tc_latency_probe(Nv.ptr(out), setmax, Nv.stream_ptr()))
        torch.cuda.synchronize()
    print("setmaxnreg %d" % setmax)
    for n, v in zip(names, out.cpu().tolist()):
        print("  %-52s %6d cycles" % (n, v))
