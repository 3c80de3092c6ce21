#!/bin/bash
# round 2 (second session), 4 GPUs: the full default bench command exactly as the driver launches it
mkdir -p gpurun_out
P=gpurun_out/round2_N4
timeout -k 5 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/round2_N4_bench.json").read().strip().splitlines()[-1])
    print("main", d["value"], d["ms_per_step"], d["roofline"]["frac"], (d.get("parity_n") or {}).get("ok"), d.get("collectives"), d.get("parity_ok"))
    for k in ("dense_path", "c3", "c5", "c4"):
        x = d.get(k) or {}
        print(k, x.get("value"), x.get("ms_per_step"), x.get("scaling"), x.get("collectives"), x.get("failed"), (x.get("parity") or {}).get("ok"), x.get("wall_s"))
except Exception as e:
    print("no line:", e)
PY
tail -3 ${P}_bench.err | cut -c1-300
