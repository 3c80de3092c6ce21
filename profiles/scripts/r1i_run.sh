mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_forest_thread --launch-skip 3 --launch-count 1 -f -o gpurun_out/r1i_forest python bench.py --workload forest --steps 1 --warmup 3 > gpurun_out/r1i_ncu_forest.log 2>&1
(timeout 900 python -m pytest tests/test_train_gpu.py tests/test_sparse_gpu.py -x -q 2>&1 | tail -5) > gpurun_out/r1i_tests.log
timeout 300 python bench.py --workload hmm --no-sparse-leg > gpurun_out/r1i_bench_hmm.json 2> gpurun_out/r1i_bench_hmm.err
cat gpurun_out/r1i_tests.log; head -c 400 gpurun_out/r1i_bench_hmm.json; tail -3 gpurun_out/r1i_ncu_forest.log
