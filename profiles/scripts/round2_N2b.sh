#!/bin/bash
# round 2 (second session), 2 GPUs, second attempt: forest + sharded Gibbs legs under torchrun, then the 2-GPU tests
mkdir -p gpurun_out
P=gpurun_out/round2_N2b
timeout -k 5 270 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --legs c5,c4 --steps 10 > ${P}_bench.json 2> ${P}_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/round2_N2b_bench.json").read().strip().splitlines()[-1])
    print("main", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("parity_n"), d.get("collectives"))
    for k in ("c5", "c4"):
        x = d.get(k) or {}
        print(k, x.get("value"), x.get("ms_per_step"), x.get("scaling"), x.get("collectives"), x.get("failed"), (x.get("parity") or {}).get("ok"), x.get("wall_s"))
except Exception as e:
    print("no line:", e)
PY
tail -3 ${P}_bench.err | cut -c1-300
timeout -k 5 280 python -m pytest tests/test_gibbs_gpu.py tests/test_forest_gpu.py -m gpu -q -k "two_gpus" > ${P}_tests.log 2>&1
echo "2-GPU tests rc=$?"; tail -6 ${P}_tests.log
