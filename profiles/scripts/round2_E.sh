#!/bin/bash
# round 2 (second session), call E: cyclic lattices, device-build through the job API, k_forest_level launch shapes
mkdir -p gpurun_out
P=gpurun_out/round2_E
timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_device_build_gpu.py -m gpu -x -q -k "cyclic or device_built" > ${P}_tests.log 2>&1
echo "tests rc=$?"; tail -12 ${P}_tests.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload forest --steps 10 > ${P}_$name.json 2> ${P}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("${P}_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$name", "ms/step %.3f" % d["ms_per_step"], "kernel_ms %.3f" % r["kernel_ms"], "frac %.3f" % r["frac"], d["parity"].get("max_rel"), d["layout"]["level_tiles"])
except Exception as e:
    print("$name failed", e)
PY
}
run v0 CB200_NO_CPU=1
run v1 CB200_NO_CPU=1 CML_FOREST_LEVEL_VARIANT=1 CML_FOREST_LEVEL_SMEM_KB=71
run v2 CB200_NO_CPU=1 CML_FOREST_LEVEL_VARIANT=2 CML_FOREST_LEVEL_SMEM_KB=52
run v3 CB200_NO_CPU=1 CML_FOREST_LEVEL_VARIANT=3 CML_FOREST_LEVEL_SMEM_KB=33
run v3pf1 CB200_NO_CPU=1 CML_FOREST_LEVEL_VARIANT=3 CML_FOREST_LEVEL_SMEM_KB=33 CML_FOREST_LEVEL_PREFETCH=1
