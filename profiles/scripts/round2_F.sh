#!/bin/bash
# round 2 (second session), call F: k_forest_level launch shapes with many small CTAs per SM
mkdir -p gpurun_out
P=gpurun_out/round2_F
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload forest --steps 10 > ${P}_$name.json 2> ${P}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("${P}_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$name", "ms/step %.3f" % d["ms_per_step"], "kernel_ms %.3f" % r["kernel_ms"], "frac %.3f" % r["frac"], d["parity"].get("max_rel"), d["layout"]["level_tiles"], d["layout"]["level_forests"])
except Exception as e:
    print("$name failed", e)
PY
}
run v3 CB200_NO_CPU=1 CML_FOREST_LEVEL_VARIANT=3 CML_FOREST_LEVEL_SMEM_KB=36
run v5 CB200_NO_CPU=1 CML_FOREST_LEVEL_VARIANT=5 CML_FOREST_LEVEL_SMEM_KB=27
run v4 CB200_NO_CPU=1 CML_FOREST_LEVEL_VARIANT=4 CML_FOREST_LEVEL_SMEM_KB=17
run v6 CB200_NO_CPU=1 CML_FOREST_LEVEL_VARIANT=6 CML_FOREST_LEVEL_SMEM_KB=13
run v7 CB200_NO_CPU=1 CML_FOREST_LEVEL_VARIANT=7 CML_FOREST_LEVEL_SMEM_KB=8
run v4pf1 CB200_NO_CPU=1 CML_FOREST_LEVEL_VARIANT=4 CML_FOREST_LEVEL_SMEM_KB=17 CML_FOREST_LEVEL_PREFETCH=1
