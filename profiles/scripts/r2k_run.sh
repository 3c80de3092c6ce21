mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -q -n 4 2>&1 | tail -8) > gpurun_out/r2k_tests.log
cat gpurun_out/r2k_tests.log
(timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > gpurun_out/r2k_smoke.log
cat gpurun_out/r2k_smoke.log
