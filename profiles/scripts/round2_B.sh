#!/bin/bash
# round 2 (second session), call B: k_forest_level -- forest parity tests over all layouts, bench of configs[4] with knob variants, ncu capture
mkdir -p gpurun_out
P=gpurun_out/round2_B
timeout 900 python -m pytest tests/test_forest_gpu.py -m gpu -x -q > ${P}_tests.log 2>&1
echo "tests rc=$?"; tail -5 ${P}_tests.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload forest --steps 10 > ${P}_$name.json 2> ${P}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("${P}_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$name", "ms/step %.3f" % d["ms_per_step"], "kernel_ms %.3f" % r["kernel_ms"], "frac %.3f" % r["frac"], d["parity"].get("max_rel"), d["layout"])
except Exception as e:
    print("$name failed", e)
PY
}
run default
run pf0 CB200_NO_CPU=1 CML_FOREST_LEVEL_PREFETCH=0
run pf4 CB200_NO_CPU=1 CML_FOREST_LEVEL_PREFETCH=4
run smem64 CB200_NO_CPU=1 CML_FOREST_LEVEL_SMEM_KB=64
run smem48 CB200_NO_CPU=1 CML_FOREST_LEVEL_SMEM_KB=48
run thr256 CB200_NO_CPU=1 CML_FOREST_LEVEL_THREADS=256
run thr256s48 CB200_NO_CPU=1 CML_FOREST_LEVEL_THREADS=256 CML_FOREST_LEVEL_SMEM_KB=48
CB200_NO_CPU=1 timeout 300 python bench.py --workload forest --steps 5 --precision 64 > ${P}_f64.json 2> ${P}_f64.err; tail -c 600 ${P}_f64.json
CB200_NO_CPU=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_forest_level --launch-skip 3 -c 1 -f -o ${P}_k_forest_level \
  python bench.py --workload forest --steps 2 > ${P}_ncu.log 2>&1
echo "ncu rc=$?"
