mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_train_gpu.py -q -k "restarts" 2>&1 | tail -30) > gpurun_out/r2i_tests.log
cat gpurun_out/r2i_tests.log
