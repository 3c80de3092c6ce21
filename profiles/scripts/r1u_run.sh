mkdir -p gpurun_out
export CML_SPARSE_MIN_SEQ=1
run() { echo "=== $*" ; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 5 "$@" 2>&1 | tail -6; echo "rc=$?"; }
(
run python -m pytest tests/test_dense_gpu.py -q -x -k "test_dense_em_matches_oracle and case1 or estep_equals or zero_probability"
run python -m pytest tests/test_sparse_gpu.py -q -x -k "estep_equals or (matches_oracle and 4-mode0)"
run python -m pytest tests/test_lane_gpu.py -q -x -k "layout_is_used or (hmm_matches and 0-mode0)"
run python -m pytest tests/test_forest_gpu.py -q -x -k "deep_random or without_stacks"
run python -m pytest tests/test_gibbs_gpu.py -q -x -k "cipher_crp_batched"
) > gpurun_out/r1u_sanitizer.log 2>&1
tail -60 gpurun_out/r1u_sanitizer.log
