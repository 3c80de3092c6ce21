#!/bin/bash
# round 2, call R: forest Gibbs tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forest_gpu.py -m gpu -x -q -k "gibbs" > gpurun_out/r2r_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2r_tests.log
tail -30 gpurun_out/r2r_tests.log
