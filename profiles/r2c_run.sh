#!/bin/bash
# round 2, call C: lane kernel with shared-memory hot tables + persistent warps; wide kernel with the cheaper stream window
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lane_gpu.py tests/test_estep_gpu.py tests/test_sparse_gpu.py -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -4 gpurun_out/r2c_tests.log
for hot in 512 256 0; do
CML_LANE_SMEM_HOT=$hot timeout 300 python bench.py --workload hmm --no-dense --steps 10 > gpurun_out/r2c_hmm_hot$hot.json 2> gpurun_out/r2c_hmm_hot$hot.err
done
timeout 300 python bench.py --workload cipher --no-dense --steps 10 > gpurun_out/r2c_cipher.json 2> gpurun_out/r2c_cipher.err
python - <<'PY'
import json
for f in ("r2c_hmm_hot512","r2c_hmm_hot256","r2c_hmm_hot0","r2c_cipher"):
    try:
        j=json.load(open(f"gpurun_out/{f}.json")); print(f, j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["ms_per_step"])
    except Exception as e: print(f, "failed", e)
PY
