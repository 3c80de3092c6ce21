#!/bin/bash
# round 2, call N (2 GPUs): NCCL inside the library: bench at N=2 (parity_n), carmel-b200 --gpus=2 test
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -q -k "two_gpus" > gpurun_out/r2n_test2gpu.log 2>&1
tail -5 gpurun_out/r2n_test2gpu.log
START=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err
echo "bench rc=$? wall $(( $(date +%s) - START ))s"
tail -5 gpurun_out/r2n_bench_n2.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2n_bench_n2.json").read().strip().splitlines()[-1])
print("main", j["value"], j["ms_per_step"], j["roofline"]["frac"], j.get("parity"), j.get("parity_n"), "coll", j.get("collectives"))
for k in ("dense_path","c3","c5"):
    d=j.get(k) or {}
    print(k, d.get("value"), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("failed"), d.get("wall_s"))
print("c3 dense", (j.get("c3") or {}).get("dense_path",{}).get("value"))
PY
