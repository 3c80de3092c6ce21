mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_lane_gpu.py -x -q 2>&1 | tail -30) > gpurun_out/r1d_lane_tests.log
timeout 300 python bench.py --workload hmm > gpurun_out/r1d_bench_hmm.json 2> gpurun_out/r1d_bench_hmm.err
CML_BENCH_NO_COUNTS=1 timeout 300 python bench.py --workload hmm --steps 5 > gpurun_out/r1d_bench_hmm_nocounts.json 2> /dev/null
timeout 300 python bench.py --workload hmm --precision 32 > gpurun_out/r1d_bench_hmm32.json 2> /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fb_lane --launch-skip 3 --launch-count 1 -f -o gpurun_out/r1d_hmm_lane python bench.py --workload hmm --steps 1 --warmup 3 > gpurun_out/r1d_ncu_hmm.log 2>&1
cat gpurun_out/r1d_lane_tests.log; head -c 3000 gpurun_out/r1d_bench_hmm.json; tail -3 gpurun_out/r1d_bench_hmm.err
