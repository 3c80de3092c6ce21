#!/bin/bash
# round 2 (second session), 2 GPUs: the bench line under torchrun (parity_n, c3 / c5 / c4 sharded) and the 2-GPU tests
mkdir -p gpurun_out
P=gpurun_out/round2_N2
timeout 900 python -m pytest tests/test_gibbs_gpu.py tests/test_forest_gpu.py tests/test_round2_gpu.py -m gpu -x -q -k "two_gpus or 2gpu or two_gpu" > ${P}_tests.log 2>&1
echo "2-GPU tests rc=$?"; tail -4 ${P}_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > ${P}_bench.json 2> ${P}_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/round2_N2_bench.json").read().strip().splitlines()[-1])
print("main", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("parity_n"), d.get("collectives"))
for k in ("dense_path", "c3", "c5", "c4"):
    x = d.get(k) or {}
    print(k, x.get("value"), x.get("ms_per_step"), x.get("scaling"), x.get("collectives"), x.get("failed"), x.get("wall_s"))
PY
tail -5 ${P}_bench.err
