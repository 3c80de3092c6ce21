#!/bin/bash
# round 2 (second session), call X: device-build tests after the input validation was added
mkdir -p gpurun_out
timeout 140 python -m pytest tests/test_device_build_gpu.py -m gpu -q -x > gpurun_out/round2_X_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/round2_X_tests.log
