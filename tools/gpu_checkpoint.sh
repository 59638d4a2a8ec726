#!/usr/bin/env bash
# One gpurun call that re-establishes the measured state of the repo (about 4 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_checkpoint.sh r02x'
# Writes gpurun_out/<tag>_*: GPU test suite, smoke, default bench (N=1) + reference arm, ncu launch list of the bench command and
# one ncu --set full capture of the CIN contraction launches of one step (both only after the plain run exited 0).
set -u
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
timeout 400 python -m pytest tests -x -q -m gpu > $out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a $out/${tag}_tests.log; tail -2 $out/${tag}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; rc=$?; echo "bench rc=$rc lines=$(wc -l < $out/${tag}_bench.json)"
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "reference arm rc=$?"
if [ "${2:-}" != "noncu" ] && [ $rc -eq 0 ]; then
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 400 --csv --log-file $out/${tag}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --profile-seconds 0.05 > $out/${tag}_ncu.log 2>&1; echo "ncu list rc=$?"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:cin_\(fwd\|bwd_dx\|bwd_dw\)_tc -s 27 -c 9 -f -o $out/${tag}_cin_full \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --profile-seconds 0.05 > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
python - <<PY
import json
for f in ("$out/${tag}_bench.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, "ms/step %.3f  frac %.3f (incl layout %.3f)  e2e %.0f  fit %s  predict %s" % (d["ms_per_step"], r["frac"], r["frac_incl_layout"],
              d["e2e"]["value"], d.get("fit_e2e", {}).get("value"), d.get("predict_e2e", {}).get("value")))
        print({k: round(v, 4) for k, v in r["other_ms_per_step"].items()})
        print({k: (round(v["achieved"]), round(v["frac"], 3)) for k, v in r["hbm_kernels"].items() if isinstance(v, dict)})
        print(d.get("cpu_baseline"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
