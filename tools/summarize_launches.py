"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share of device time.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [first_id] > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name)
    return name if len(name) < 110 else name[:107] + "..."


def main():
    path = sys.argv[1]
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr, data = rows[0], rows[1:]
    ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    tot = 0.0
    for r in data:
        if int(r[ii]) < first:
            continue
        ns = float(r[vi].replace(",", ""))
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += ns
        tot += ns
    print("| kernel | launches | total us | share | grid | block |")
    print("|---|---|---|---|---|---|")
    for k, (n, ns, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% | %s | %s |" % (k, n, ns / 1e3, 100 * ns / tot, g, b))
    print("\ntotal device time of the %d listed launches: %.3f ms" % (sum(a[0] for a in agg.values()), tot / 1e6))


if __name__ == "__main__":
    main()
