mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_forest_gpu.py -x -q -k "thread or random or sample or cipher" 2>&1 | tail -15) > gpurun_out/r1l_tests.log
timeout 400 python bench.py --workload forest > gpurun_out/r1l_bench_forest.json 2> gpurun_out/r1l_bench_forest.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_forest_thread --launch-skip 3 --launch-count 1 -f -o gpurun_out/r1l_forest python bench.py --workload forest --steps 1 --warmup 3 > gpurun_out/r1l_ncu_forest.log 2>&1
cat gpurun_out/r1l_tests.log; head -c 1500 gpurun_out/r1l_bench_forest.json; tail -3 gpurun_out/r1l_bench_forest.err
