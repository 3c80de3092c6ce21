#!/bin/bash
# round 2, call I: wide kernel variants (addressing, L2 prefetch, warps per CTA)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload cipher --no-dense --steps 20 --no-sparse-leg > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2i_$name.json")); print("$name", j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["ms_per_step"])
except Exception as e: print("$name", "failed", e)
PY
}
run base X=1
run ga CML_WIDE_GA=1
run pf12 CML_WIDE_PF=12
run pf4 CML_WIDE_PF=4
run wpc16 CML_WIDE_WPC=16
run wpc12 CML_WIDE_WPC=12
run wpc10 CML_WIDE_WPC=10
