#!/bin/bash
# round 2 (second session), call H: lane kernel with the cp.async stream ring; forest labels with hot rows; tile sizes
mkdir -p gpurun_out
P=gpurun_out/round2_H
timeout 900 python -m pytest tests/test_lane_gpu.py tests/test_round2_gpu.py -m gpu -x -q > ${P}_lane_tests.log 2>&1
echo "lane tests rc=$?"; tail -4 ${P}_lane_tests.log
timeout 600 python -m pytest tests/test_forest_gpu.py -m gpu -x -q -k "level_layout or sample_forests or random_forests or zero_probability" > ${P}_forest_tests.log 2>&1
echo "forest tests rc=$?"; tail -4 ${P}_forest_tests.log
for mb in 2 3; do
  CML_LANE_MINB=$mb timeout 300 python bench.py --workload hmm --no-dense --no-sparse-leg --steps 10 > ${P}_lane_minb$mb.json 2> ${P}_lane_minb$mb.err
  python -c "
import json; d=json.loads(open('${P}_lane_minb$mb.json').read().strip().splitlines()[-1]); print('lane minb$mb', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('parity'))"
done
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload forest --steps 10 > ${P}_$name.json 2> ${P}_$name.err
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
run kb10 CB200_NO_CPU=1 CML_FOREST_LEVEL_SMEM_KB=10
run kb8 CB200_NO_CPU=1 CML_FOREST_LEVEL_SMEM_KB=8
run kb6 CB200_NO_CPU=1 CML_FOREST_LEVEL_SMEM_KB=6
