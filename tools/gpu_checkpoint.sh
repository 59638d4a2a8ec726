#!/usr/bin/env bash
# One gpurun call that re-establishes the measured state of the repo (≈ 3 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 400 -- 'bash tools/gpu_checkpoint.sh r02a'
# Writes gpurun_out/<tag>_*.{log,json,csv}: GPU test suite, smoke, default bench (N=1), the dW geometry A/B prepared at the end of
# round 1 (profiles/r01d_cin_dw_findings.md), and the ncu launch list of the bench command (after the plain run exited 0).
set -u
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
timeout 150 python -m pytest tests -x -q -m gpu > $out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a $out/${tag}_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 120 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$? lines=$(wc -l < $out/${tag}_bench.json)"
XDFM_CIN_DW_JP=1 timeout 100 python -m pytest tests/test_gpu_tc.py tests/test_gpu_cin.py -x -q -m gpu > $out/${tag}_jp1_tests.log 2>&1; echo "jp1 tests rc=$?"
XDFM_CIN_DW_JP=1 timeout 90 python bench.py --steps 30 --no-cpu-baseline > $out/${tag}_bench_jp1.json 2> $out/${tag}_bench_jp1.err; echo "jp1 bench rc=$?"
timeout 60 python tools/bench_bag.py > $out/${tag}_bench_bag.log 2>&1; echo "bag rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
for f in ("$out/${tag}_bench.json", "$out/${tag}_bench_jp1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, "ms/step %.3f  frac %.3f  cin_bwd %.3f ms  cin_fwd %.3f ms" % (d["ms_per_step"], r["frac"], r["other_ms_per_step"]["cin_bwd"], r["other_ms_per_step"]["cin_fwd"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY
