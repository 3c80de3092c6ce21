#!/bin/bash
# round 2 (second session), call G: k_forest_level with two tile classes (default 128 x 16), f64, ncu capture
mkdir -p gpurun_out
P=gpurun_out/round2_G
timeout 600 python -m pytest tests/test_forest_gpu.py -m gpu -x -q -k "level_layout or sample_forests or random_forests or zero_probability or big_forests" > ${P}_tests.log 2>&1
echo "tests rc=$?"; tail -4 ${P}_tests.log
run() {  # name, args, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload forest --steps 10 $EXTRA > ${P}_$name.json 2> ${P}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("${P}_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$name", "ms/step %.3f" % d["ms_per_step"], "kernel_ms %.3f" % r["kernel_ms"], "frac %.3f" % r["frac"], d["parity"].get("max_rel"), d["layout"]["level_tiles"], d["layout"]["small_tiles"])
except Exception as e:
    print("$name failed", e)
PY
}
run default CB200_NO_CPU=1
run kb11 CB200_NO_CPU=1 CML_FOREST_LEVEL_SMEM_KB=11
EXTRA="--precision 64" run f64 CB200_NO_CPU=1
EXTRA="--precision 64" run f64v3 CB200_NO_CPU=1 CML_FOREST_LEVEL_VARIANT=3 CML_FOREST_LEVEL_SMEM_KB=36
CB200_NO_CPU=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_forest_level --launch-skip 3 -c 1 -f -o ${P}_k_forest_level \
  python bench.py --workload forest --steps 2 > ${P}_ncu.log 2>&1
echo "ncu rc=$?"
