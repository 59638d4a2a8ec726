"""Top source lines by warp-stall samples from an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_source_hot.py report.ncu-rep <kernel-index> [top_n]"""
import csv
import subprocess
import sys

rep, kidx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if len(row) == 2 and row[0] == "File Path":
        cur = {"file": row[1], "rows": []}
        blocks.append(cur)
    elif len(row) == 2 and row[0] == "Function Name" and cur is not None:
        cur["fn"] = row[1]
    elif cur is not None and len(row) > 10:
        if row[0] == "Line No":
            cur["hdr"] = row
        else:
            cur["rows"].append(row)
# one kernel may span several files (inlined headers): group consecutive blocks with the same function name
kernels = []
for b in blocks:
    if kernels and kernels[-1][0] == b.get("fn"):
        kernels[-1][1].append(b)
    else:
        kernels.append((b.get("fn"), [b]))
fn, bl = kernels[kidx]
print("kernel:", fn)
lines = []
total = 0
for b in bl:
    h = b["hdr"]
    si = h.index("# Samples")
    stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    for r in b["rows"]:
        if r[0] == "":
            continue
        def num(x):
            try:
                return int(float(x))
            except ValueError:
                return 0
        n = num(r[si])
        total += n
        st = sorted(((num(r[i]), nm) for i, nm in stall_cols), reverse=True)[:3]
        lines.append((n, b["file"].split("/")[-1], r[0], r[1].strip()[:90], st))
lines.sort(reverse=True)
print("total samples:", total)
for n, f, ln, src, st in lines[:top]:
    print("%6d %5.1f%%  %s:%s  %s   [%s]" % (n, 100.0 * n / max(total, 1), f, ln, src, ", ".join("%s=%d" % (nm[6:], c) for c, nm in st if c)))
