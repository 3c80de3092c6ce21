mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gibbs_gpu.py -x -q 2>&1 | tail -15) > gpurun_out/r1m_tests.log
timeout 600 python bench.py --workload gibbs > gpurun_out/r1m_bench_gibbs.json 2> gpurun_out/r1m_bench_gibbs.err
cat gpurun_out/r1m_tests.log; head -c 2500 gpurun_out/r1m_bench_gibbs.json; tail -5 gpurun_out/r1m_bench_gibbs.err
