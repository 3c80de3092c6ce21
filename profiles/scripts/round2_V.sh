#!/bin/bash
# round 2 (second session), call V: --viterbi (cml_viterbi) against the oracle
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_round2_gpu.py -m gpu -q -k "viterbi" > gpurun_out/round2_V_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/round2_V_tests.log
