#!/bin/bash
# round 2 (second session), call D: warp tiles in k_forest_level; remaining device-build tests; lane-kernel register cap; build times at 1M sentences
mkdir -p gpurun_out
P=gpurun_out/round2_D
timeout 600 python -m pytest tests/test_device_build_gpu.py -m gpu -x -q -k "cascades or auto" > ${P}_build_tests.log 2>&1
echo "build tests rc=$?"; tail -4 ${P}_build_tests.log
timeout 600 python -m pytest tests/test_forest_gpu.py -m gpu -x -q -k "sample_forests or norm_and or random_forests or zero_probability or big_forests or cipher_forests" > ${P}_forest_tests.log 2>&1
echo "forest tests rc=$?"; tail -4 ${P}_forest_tests.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload forest --steps 10 > ${P}_$name.json 2> ${P}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("${P}_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$name", "ms/step %.3f" % d["ms_per_step"], "kernel_ms %.3f" % r["kernel_ms"], "frac %.3f" % r["frac"], d["parity"].get("max_rel"), d["layout"]["level_tiles"], d["layout"]["warp_tiles"])
except Exception as e:
    print("$name failed", e)
PY
}
run default CB200_NO_CPU=1
run nowarp CB200_NO_CPU=1 CML_FOREST_LEVEL_WARP=0
run pf0 CB200_NO_CPU=1 CML_FOREST_LEVEL_PREFETCH=0
run smem64 CB200_NO_CPU=1 CML_FOREST_LEVEL_SMEM_KB=64
CB200_NO_CPU=1 timeout 300 python bench.py --workload forest --steps 5 --precision 64 > ${P}_f64.json 2> ${P}_f64.err; python -c "
import json; d=json.loads(open('${P}_f64.json').read().strip().splitlines()[-1]); print('f64', d['ms_per_step'], d['roofline']['kernel_ms'], d['parity']['max_rel'], d['layout'])"
CB200_NO_CPU=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_forest_level --launch-skip 3 -c 1 -f -o ${P}_k_forest_level \
  python bench.py --workload forest --steps 2 > ${P}_ncu.log 2>&1
echo "ncu rc=$?"
for mb in 2 3; do
  CML_LANE_MINB=$mb timeout 300 python bench.py --workload hmm --no-dense --no-sparse-leg --steps 10 > ${P}_lane_minb$mb.json 2> ${P}_lane_minb$mb.err
  python -c "
import json; d=json.loads(open('${P}_lane_minb$mb.json').read().strip().splitlines()[-1]); print('lane minb$mb', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
python - <<'PY' > ${P}_build_times.txt 2>&1
import os, subprocess, tempfile, time, sys
sys.path.insert(0, ".")
from carmel_b200 import synth, CLI_PATH
d = tempfile.mkdtemp()
w = synth.write_hmm(os.path.join(d, "h"), n_sent=1000000)
for how in ("--host-build", "--device-build"):
    t = time.time()
    p = subprocess.run([CLI_PATH, *w["argv"], "--trellis-only", how], capture_output=True, text=True)
    print("hmm1M", how, "wall %.2fs" % (time.time() - t), [l for l in p.stderr.splitlines() if "Built" in l or "Device-side" in l or "ERROR" in l])
PY
cat ${P}_build_times.txt
