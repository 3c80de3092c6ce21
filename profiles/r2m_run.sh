#!/bin/bash
# round 2, call M: the whole GPU suite (incl. tests/test_round2_gpu.py)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2m_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2m_tests.log
tail -25 gpurun_out/r2m_tests.log
