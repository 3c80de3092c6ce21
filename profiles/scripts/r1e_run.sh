mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_sparse_gpu.py tests/test_dense_gpu.py -x -q 2>&1 | tail -30) > gpurun_out/r1e_tests.log
timeout 300 python bench.py --workload hmm > gpurun_out/r1e_bench_hmm.json 2> gpurun_out/r1e_bench_hmm.err
timeout 300 python bench.py --workload hmm --precision 32 --no-sparse-leg > gpurun_out/r1e_bench_hmm32.json 2> /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fb_sparse --launch-skip 3 --launch-count 1 -f -o gpurun_out/r1e_hmm_sparse python bench.py --workload hmm --steps 1 --warmup 3 --no-sparse-leg > gpurun_out/r1e_ncu_hmm.log 2>&1
cat gpurun_out/r1e_tests.log; head -c 2500 gpurun_out/r1e_bench_hmm.json; tail -3 gpurun_out/r1e_bench_hmm.err
