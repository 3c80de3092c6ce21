mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2c_smoke.log 2>&1
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r2c_tests.log
timeout 300 python bench.py > gpurun_out/r2c_bench_cipher.json 2> gpurun_out/r2c_bench_cipher.err
tail -3 gpurun_out/r2c_smoke.log; cat gpurun_out/r2c_tests.log; head -c 400 gpurun_out/r2c_bench_cipher.json
