mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_dense_gpu.py -q -k "long_lines" 2>&1 | tail -4) > gpurun_out/r1w_tests.log
for n in 2 4 8; do CML_SPARSE_XI_TABLES=$n timeout 300 python bench.py --workload hmm --no-sparse-leg --steps 10 > gpurun_out/r1w_hmm_xi$n.json 2>/dev/null; done
cat gpurun_out/r1w_tests.log; python - <<'PY'
import json
for n in (2,4,8):
    j=json.loads(open(f"gpurun_out/r1w_hmm_xi{n}.json").read().strip().splitlines()[-1])
    print("xi tables", n, "%.3g"%j["value"], "step %.3f"%j["ms_per_step"], "kernel %.4f"%j["roofline"]["kernel_ms"])
PY
