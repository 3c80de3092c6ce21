mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_dense_gpu.py -x -q 2>&1 | tail -30) > gpurun_out/r1c_dense_tests.log
timeout 300 python bench.py > gpurun_out/r1c_bench_cipher.json 2> gpurun_out/r1c_bench_cipher.err
timeout 300 python bench.py --precision 32 > gpurun_out/r1c_bench_cipher32.json 2> gpurun_out/r1c_bench_cipher32.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fb_dense --launch-skip 3 --launch-count 1 -f -o gpurun_out/r1c_dense python bench.py --steps 1 --warmup 3 --no-sparse-leg > gpurun_out/r1c_ncu_dense.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fb_ell --launch-skip 3 --launch-count 1 -f -o gpurun_out/r1c_hmm python bench.py --workload hmm --steps 1 --warmup 3 > gpurun_out/r1c_ncu_hmm.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_cipher_launches.csv python bench.py --steps 2 --warmup 3 --no-sparse-leg > gpurun_out/r1c_launches.log 2>&1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r1c_tests.log
cat gpurun_out/r1c_dense_tests.log gpurun_out/r1c_tests.log; cat gpurun_out/r1c_bench_cipher.json | head -c 6000; tail -3 gpurun_out/r1c_bench_cipher.err
