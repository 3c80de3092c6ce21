mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_sparse_gpu.py tests/test_abi_cpu.py -q 2>&1 | tail -3) > gpurun_out/r2g_tests.log
cat gpurun_out/r2g_tests.log
