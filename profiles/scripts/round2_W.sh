#!/bin/bash
# round 2 (second session), call W: the forest layouts and the E-step parity tests on the final library
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_forest_gpu.py -m gpu -q -x -k "sample_forests or norm_and or zero_probability or big_forests" > gpurun_out/round2_W_forest.log 2>&1
echo "forest rc=$?"; tail -2 gpurun_out/round2_W_forest.log
timeout 110 python -m pytest tests/test_estep_gpu.py tests/test_lane_gpu.py -m gpu -q -x > gpurun_out/round2_W_estep.log 2>&1
echo "estep rc=$?"; tail -2 gpurun_out/round2_W_estep.log
