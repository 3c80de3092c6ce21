mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_dense_gpu.py -q -k "tensor_core or synthetic_27" 2>&1 | tail -6) > gpurun_out/r2a_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30 -c 30 --csv --log-file gpurun_out/r2a_tc_launches.csv python bench.py --scale 64 --no-sparse-leg --steps 3 --precision 32 > gpurun_out/r2a_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_tc --launch-skip 6 --launch-count 2 -f -o gpurun_out/r2a_dense_tc python bench.py --scale 64 --no-sparse-leg --steps 1 --precision 32 > gpurun_out/r2a_ncu.log 2>&1
cat gpurun_out/r2a_tests.log; grep -c k_dense_tc gpurun_out/r2a_tc_launches.csv
