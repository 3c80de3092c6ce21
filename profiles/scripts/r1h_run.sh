mkdir -p gpurun_out
timeout 300 python bench.py --workload hmm > gpurun_out/r1h_bench_hmm.json 2> gpurun_out/r1h_bench_hmm.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1h_hmm_launches.csv python bench.py --workload hmm --steps 2 --warmup 3 --no-sparse-leg > gpurun_out/r1h_launches.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r1h_bench_cipher_n2.json 2> gpurun_out/r1h_bench_cipher_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload hmm > gpurun_out/r1h_bench_hmm_n2.json 2> gpurun_out/r1h_bench_hmm_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --workload forest > gpurun_out/r1h_bench_forest_n2.json 2> gpurun_out/r1h_bench_forest_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --workload gibbs > gpurun_out/r1h_bench_gibbs_n2.json 2> gpurun_out/r1h_bench_gibbs_n2.err
for f in gpurun_out/r1h_bench_*.json; do echo $f; head -c 600 $f; echo; done; tail -5 gpurun_out/r1h_bench_*_n2.err
