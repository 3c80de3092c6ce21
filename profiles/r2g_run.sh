#!/bin/bash
# round 2, call G: wide kernel with aligned shared addressing (10 instructions per row)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_estep_gpu.py tests/test_train_gpu.py tests/test_dense_gpu.py -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2g_tests.log
tail -3 gpurun_out/r2g_tests.log
timeout 300 python bench.py --workload cipher --no-dense --steps 20 > gpurun_out/r2g_cipher.json 2> gpurun_out/r2g_cipher.err
timeout 300 python bench.py --workload cipher --no-dense --steps 20 --precision 32 > gpurun_out/r2g_cipher32.json 2> gpurun_out/r2g_cipher32.err
python - <<'PY'
import json
for f in ("r2g_cipher","r2g_cipher32"):
    try:
        j=json.load(open(f"gpurun_out/{f}.json")); print(f, j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["ms_per_step"])
    except Exception as e: print(f, "failed", e)
PY
