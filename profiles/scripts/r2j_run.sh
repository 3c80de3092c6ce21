mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_forest_gpu.py -q -k "restarts or sample_forests or errors" 2>&1 | tail -30) > gpurun_out/r2j_tests.log
cat gpurun_out/r2j_tests.log
