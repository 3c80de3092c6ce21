#!/bin/bash
# round 2, call D: lane kernel with the bulk-copy stream pipeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lane_gpu.py tests/test_estep_gpu.py -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2d_tests.log
tail -4 gpurun_out/r2d_tests.log
timeout 300 python bench.py --workload hmm --no-dense --steps 10 > gpurun_out/r2d_hmm.json 2> gpurun_out/r2d_hmm.err
CML_BENCH_NO_COUNTS=1 timeout 300 python bench.py --workload hmm --no-dense --steps 10 > gpurun_out/r2d_hmm_nocounts.json 2> gpurun_out/r2d_hmm_nocounts.err
python - <<'PY'
import json
for f in ("r2d_hmm","r2d_hmm_nocounts"):
    try:
        j=json.load(open(f"gpurun_out/{f}.json")); print(f, j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["ms_per_step"])
    except Exception as e: print(f, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fb_lane -s 3 -c 1 -o gpurun_out/r2d_k_fb_lane_f64 python bench.py --workload hmm --no-dense --steps 2 --warmup 3 > gpurun_out/r2d_ncu_lane.log 2>&1
