#!/bin/bash
# round 2, call H: wide kernel + L2 prefetch ahead of the bulk-copy ring
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_estep_gpu.py -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2h_tests.log
tail -3 gpurun_out/r2h_tests.log
timeout 300 python bench.py --workload cipher --no-dense --steps 20 > gpurun_out/r2h_cipher.json 2> gpurun_out/r2h_cipher.err
python - <<'PY'
import json
for f in ("r2h_cipher",):
    try:
        j=json.load(open(f"gpurun_out/{f}.json")); print(f, j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["ms_per_step"])
    except Exception as e: print(f, "failed", e)
PY
