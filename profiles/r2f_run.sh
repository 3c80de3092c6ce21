#!/bin/bash
# round 2, call F: wide kernel profile after the cheaper stream window; restored lane kernel check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lane_gpu.py tests/test_estep_gpu.py -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2f_tests.log
tail -3 gpurun_out/r2f_tests.log
timeout 300 python bench.py --workload hmm --no-dense --steps 10 > gpurun_out/r2f_hmm.json 2> gpurun_out/r2f_hmm.err
timeout 300 python bench.py --workload cipher --no-dense --steps 10 > gpurun_out/r2f_cipher.json 2> gpurun_out/r2f_cipher.err
python - <<'PY'
import json
for f in ("r2f_hmm","r2f_cipher"):
    try:
        j=json.load(open(f"gpurun_out/{f}.json")); print(f, j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["ms_per_step"])
    except Exception as e: print(f, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fb_wide -s 3 -c 1 -o gpurun_out/r2f_k_fb_wide_f64 python bench.py --workload cipher --no-dense --steps 2 --warmup 3 > gpurun_out/r2f_ncu_wide.log 2>&1
