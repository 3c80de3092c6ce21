#!/bin/bash
# round 2, call S: --expectation tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gibbs_gpu.py -m gpu -x -q -k "expectation" > gpurun_out/r2s_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2s_tests.log
tail -30 gpurun_out/r2s_tests.log
