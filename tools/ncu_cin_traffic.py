"""Per-launch summary of an `ncu --set full` capture of bench.py (CIN kernels): duration, DRAM bytes, tensor-pipe activity.
Writes profiles/<tag>_ncu_cin_traffic.json (read by bench.py for `roofline.traffic`) and a markdown table.

    python tools/ncu_cin_traffic.py gpurun_out/r02_cin_full.ncu-rep r02 [launches_per_step]

The capture: ncu --set full --clock-control none --import-source on -k regex:cin_(fwd|bwd)_ -s <skip> -c <n> python bench.py ..."""
import csv
import datetime
import glob
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
per_step = int(sys.argv[3]) if len(sys.argv) > 3 else 9
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}


def find(name):
    for n, i in col.items():
        if n == name:
            return i
    return None


want = {"dur_us": "gpu__time_duration.sum", "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
        "tensor_pct": "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "tensor_pipe_cycles_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "tmem_pipe_pct": "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "regs": "launch__registers_per_thread", "sm_mhz": "sm__cycles_elapsed.avg.per_second"}
units = rows[1]
kernels = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    k = {"name": r[col["Kernel Name"]][:70]}
    for key, metric in want.items():
        i = find(metric)
        if i is None or r[i] == "":
            k[key] = None
            continue
        v = float(r[i].replace(",", ""))
        u = units[i]
        if key in ("dram_rd", "dram_wr"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        if key == "dur_us":
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
        k[key] = v
    kernels.append(k)
cin = [k for k in kernels if "cin_fwd_tc" in k["name"] or "cin_bwd_dx_tc" in k["name"] or "cin_bwd_dw_tc" in k["name"]]
step = cin[:per_step]
srcs = sorted(glob.glob(os.path.join(ROOT, "xdeepfm-pytorch_b200", "csrc", "*")))
h = hashlib.sha256()
for p in srcs:
    h.update(open(p, "rb").read())
total = sum((k["dram_rd"] or 0) + (k["dram_wr"] or 0) for k in step)
summary = {"dram_bytes_per_step": total, "launches_per_step": len(step), "csrc_hash": h.hexdigest()[:16],
           "when": datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%MZ"), "report": os.path.basename(rep), "launches": step}
json.dump(summary, open(os.path.join(ROOT, "profiles", "%s_ncu_cin_traffic.json" % tag), "w"), indent=1)
with open(os.path.join(ROOT, "profiles", "%s_ncu_cin_full.md" % tag), "w") as f:
    f.write("# ncu --set full, CIN contraction launches of one cfg2 step (%s, csrc %s)\n\n" % (os.path.basename(rep), summary["csrc_hash"]))
    f.write("| kernel | us | DRAM read MB | DRAM write MB | tensor pipe active % | tensor inst % | L2 % | DRAM % | regs |\n|---|---|---|---|---|---|---|---|---|\n")
    for k in step:
        f.write("| `%s` | %.1f | %.1f | %.1f | %s | %s | %s | %s | %s |\n" % (
            k["name"], k["dur_us"] or 0, (k["dram_rd"] or 0) / 1e6, (k["dram_wr"] or 0) / 1e6, k["tensor_pipe_cycles_pct"], k["tensor_pct"],
            k["l2_pct"], k["dram_pct"], k["regs"]))
    f.write("\nDRAM bytes of the %d launches: %.1f MB\n" % (len(step), total / 1e6))
print(json.dumps({k: v for k, v in summary.items() if k != "launches"}))
