mkdir -p gpurun_out
timeout 600 python bench.py --scale 64 --no-sparse-leg --steps 10 --precision 32 --unlock-lm > gpurun_out/r2f_unlocked_tc.json 2>/dev/null
CML_DENSE_TC=0 timeout 600 python bench.py --scale 64 --no-sparse-leg --steps 10 --precision 32 --unlock-lm > gpurun_out/r2f_unlocked_fma.json 2>/dev/null
(timeout 300 python -m pytest tests/test_dense_gpu.py -q -k tensor_core 2>&1 | tail -3) > gpurun_out/r2f_tests.log
cat gpurun_out/r2f_tests.log; python - <<'PY'
import json
for f in ("unlocked_tc","unlocked_fma"):
    j=json.loads(open(f"gpurun_out/r2f_{f}.json").read().strip().splitlines()[-1])
    print(f, "%.3g"%j["value"], "step %.3f"%j["ms_per_step"], "kernel %.4f"%j["roofline"]["kernel_ms"], j["roofline"]["kernel"][:30], j["layout"])
PY
