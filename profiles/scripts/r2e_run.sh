mkdir -p gpurun_out
(CML_DENSE_TC=1 timeout 900 python -m pytest tests/test_dense_gpu.py -q 2>&1 | tail -25) > gpurun_out/r2e_tests.log
cat gpurun_out/r2e_tests.log | cut -c1-300
