#!/bin/bash
# round 2, call J: configs[2] at its full size (1M sentences) on one GPU: build time, memory, both paths
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/r2j_mem.txt; nproc >> gpurun_out/r2j_mem.txt
timeout 1200 python bench.py --workload hmm --scale 8 --steps 5 > gpurun_out/r2j_hmm_1m.json 2> gpurun_out/r2j_hmm_1m.err
tail -5 gpurun_out/r2j_hmm_1m.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r2j_hmm_1m.json"))
print("dense", j["value"], j["ms_per_step"], j["e2e"]["lattices"])
s=j["sparse_path"]; print("sparse", s["value"], s["ms_per_step"], s["roofline"]["frac"], s["e2e"]["lattices"], j["totals"])
PY
cat gpurun_out/r2j_mem.txt
