mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 40 --csv --log-file gpurun_out/r1n_gibbs_launches.csv python bench.py --workload gibbs --steps 3 --warmup 3 --no-dense-leg > gpurun_out/r1n_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gibbs_dense$ --launch-skip 3 --launch-count 1 -f -o gpurun_out/r1n_gibbs_dense python bench.py --workload gibbs --steps 1 --warmup 3 > gpurun_out/r1n_ncu.log 2>&1
tail -3 gpurun_out/r1n_ncu.log
