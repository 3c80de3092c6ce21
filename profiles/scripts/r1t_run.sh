mkdir -p gpurun_out
p=29600
for w in cipher hmm forest gibbs; do
  p=$((p+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 2 --steps 10 --warmup 3 --workload $w > gpurun_out/r1t_bench_${w}_n2.json 2> gpurun_out/r1t_bench_${w}_n2.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/r1t_bench_reference_n2.json 2>&1
python - <<'PY'
import json
for w in ("cipher","hmm","forest","gibbs"):
    try:
        j=json.loads(open(f"gpurun_out/r1t_bench_{w}_n2.json").read().strip().splitlines()[-1])
        print(w, "N=%d"%j["n_gpus"], "%.3g"%j["value"], "step %.3f"%j["ms_per_step"])
    except Exception as e:
        print(w, "ERR", e); print(open(f"gpurun_out/r1t_bench_{w}_n2.err").read()[-800:])
print(open("gpurun_out/r1t_bench_reference_n2.json").read()[-300:])
PY
