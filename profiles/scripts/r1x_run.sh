mkdir -p gpurun_out
for n in 2 3 4; do CML_SPARSE_XI_TABLES=$n timeout 300 python bench.py --workload hmm --no-sparse-leg --steps 10 > gpurun_out/r1x_hmm_xi$n.json 2>/dev/null; done
(timeout 600 python -m pytest tests/test_sparse_gpu.py -q 2>&1 | tail -3) > gpurun_out/r1x_tests.log
cat gpurun_out/r1x_tests.log; python - <<'PY'
import json
for n in (2,3,4):
    j=json.loads(open(f"gpurun_out/r1x_hmm_xi{n}.json").read().strip().splitlines()[-1])
    print("xi tables", n, "%.3g"%j["value"], "step %.3f"%j["ms_per_step"], "kernel %.4f"%j["roofline"]["kernel_ms"])
PY
