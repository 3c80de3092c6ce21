#!/bin/bash
# round 2, call L: new round-2 tests, then the whole GPU suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_round2_gpu.py -m gpu -x -q --durations=15 > gpurun_out/r2l_round2.log 2>&1
echo "round2 rc=$?" >> gpurun_out/r2l_round2.log
tail -40 gpurun_out/r2l_round2.log
