"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMA use (cuobjdump -sass of the built library).
usage: python tools/sass_table.py > profiles/sass_tcgen05.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "xdeepfm-pytorch_b200", "libxdfm_sm100a.so")
MNEMONICS = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "USETMAXREG", "FFMA2"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = re.sub(r"^void ", "", cur)
        counts.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    for mn in MNEMONICS:
        if re.search(r"\b%s\b|\b%s\." % (mn, mn), line):
            counts[cur][mn] += 1
stamp = open(os.path.join(ROOT, "xdeepfm-pytorch_b200", "build", "stamp")).read()[:16]
print("# SASS evidence of tcgen05 / TMA per kernel (`cuobjdump -sass libxdfm_sm100a.so`, build stamp %s)\n" % stamp)
print("UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, UBLKCP = bulk copy, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,")
print("SYNCS = mbarrier, USETMAXREG = setmaxnreg, FFMA2 = fma.rn.f32x2.  Kernels without any of them are omitted.\n")
print("| kernel | " + " | ".join(MNEMONICS) + " |")
print("|---|" + "---|" * len(MNEMONICS))
for k, c in counts.items():
    if sum(c[m] for m in MNEMONICS[:6]) == 0:
        continue
    print("| `%s` | " % k + " | ".join(str(c[m]) for m in MNEMONICS) + " |")
print("\nlibrary size: %.1f MB" % (os.path.getsize(LIB) / 1e6))
