#!/bin/bash
# round 2, call Q: new forest tests (checkpoints, viterbi), fem export test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forest_gpu.py tests/test_round2_gpu.py -m gpu -x -q -k "checkpoint or viterbi or fem_export or two_gpus" > gpurun_out/r2q_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2q_tests.log
tail -30 gpurun_out/r2q_tests.log
