#!/bin/bash
# round 2, call B: ncu captures of k_fb_lane (hmm) and k_fb_wide (cipher), no-count floors
mkdir -p gpurun_out
CML_BENCH_NO_COUNTS=1 timeout 300 python bench.py --workload hmm --no-dense --steps 10 > gpurun_out/r2b_hmm_nocounts.json 2> gpurun_out/r2b_hmm_nocounts.err
python - <<'PY'
import json
for f in ("r2b_hmm_nocounts",):
    j=json.load(open(f"gpurun_out/{f}.json")); print(f, j["roofline"]["kernel_ms"], j["roofline"]["frac"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fb_lane -s 3 -c 1 -o gpurun_out/r2b_k_fb_lane_f64 python bench.py --workload hmm --no-dense --steps 2 --warmup 3 > gpurun_out/r2b_ncu_lane.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fb_wide -s 3 -c 1 -o gpurun_out/r2b_k_fb_wide_f64 python bench.py --workload cipher --no-dense --steps 2 --warmup 3 > gpurun_out/r2b_ncu_wide.log 2>&1
CML_BENCH_NO_COUNTS=1 timeout 300 python bench.py --workload cipher --no-dense --steps 10 > gpurun_out/r2b_cipher_nocounts.json 2> gpurun_out/r2b_cipher_nocounts.err
ls -la gpurun_out/
