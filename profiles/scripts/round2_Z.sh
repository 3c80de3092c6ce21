#!/bin/bash
# round 2 (second session), call Z: validation of the tree as committed -- full GPU test suite, smoke, the default bench
# line (all legs), the reference arm, the ncu launch list of the bench command, compute-sanitizer on the new kernels
mkdir -p gpurun_out
P=gpurun_out/round2_Z
timeout 1700 python -m pytest tests -m gpu -q > ${P}_tests.log 2>&1
echo "tests rc=$?" | tee -a ${P}_tests.log; tail -6 ${P}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 ${P}_smoke.log
timeout 900 python bench.py > ${P}_bench_n1.json 2> ${P}_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/round2_Z_bench_n1.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("main", d["value"], d["ms_per_step"], r["kernel"][:20], r["frac"], "e2e", d["e2e"]["value"], "parity", d["parity"]["ok"], d.get("parity_ok"))
for k in ("dense_path", "c3", "c5", "c4"):
    x = d.get(k) or {}
    print(k, x.get("value"), x.get("ms_per_step"), (x.get("roofline") or {}).get("frac"), (x.get("parity") or {}).get("ok"), x.get("wall_s"))
print("cli", d.get("e2e_cli"))
PY
timeout 300 python bench.py --impl reference --steps 2 > ${P}_bench_reference.json 2> ${P}_bench_reference.err; echo "reference rc=$?"; tail -c 400 ${P}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${P}_launches.csv \
  python bench.py --steps 2 --warmup 3 --legs none > ${P}_launches_bench.log 2>&1
echo "launch list rc=$?"
for t in "tests/test_forest_gpu.py -k sample_forests" "tests/test_device_build_gpu.py -k reference_fixtures" "tests/test_round2_gpu.py -k cyclic"; do
  timeout 600 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 python -m pytest $t -m gpu -x -q >> ${P}_sanitizer.log 2>&1
  echo "memcheck [$t] rc=$?" | tee -a ${P}_sanitizer.log
done
grep -E "ERROR SUMMARY|passed|failed" ${P}_sanitizer.log | tail -12
