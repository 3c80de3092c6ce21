#!/bin/bash
# round 2, call A: GPU tests of the factored wide / lane kernels + sparse-path bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 300 python bench.py --workload cipher --no-dense --steps 10 > gpurun_out/r2a_bench_cipher_sparse.json 2> gpurun_out/r2a_bench_cipher_sparse.err
tail -c 1500 gpurun_out/r2a_bench_cipher_sparse.json
timeout 300 python bench.py --workload hmm --no-dense --steps 10 > gpurun_out/r2a_bench_hmm_sparse.json 2> gpurun_out/r2a_bench_hmm_sparse.err
tail -c 1500 gpurun_out/r2a_bench_hmm_sparse.json
