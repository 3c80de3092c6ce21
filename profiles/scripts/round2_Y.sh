#!/bin/bash
# round 2 (second session), call Y: the bench line of the final tree (after the bench.py refactor: finish_line, CPU side
# channel), one GPU, all legs; GPU tests touched since call Z
mkdir -p gpurun_out
P=gpurun_out/round2_Y
timeout 700 python bench.py > ${P}_bench_n1.json 2> ${P}_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/round2_Y_bench_n1.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("main", d["value"], d["ms_per_step"], r["frac"], "e2e", d["e2e"]["value"], "parity_ok", d.get("parity_ok"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
for k in ("dense_path", "c3", "c5", "c4"):
    x = d.get(k) or {}
    print(k, x.get("value"), x.get("ms_per_step"), (x.get("roofline") or {}).get("frac"), (x.get("parity") or {}).get("ok"), x.get("failed"), x.get("wall_s"))
PY
timeout 400 python -m pytest tests/test_gibbs_gpu.py tests/test_round2_gpu.py -m gpu -q -x > ${P}_tests.log 2>&1; echo "tests rc=$?"; tail -3 ${P}_tests.log
