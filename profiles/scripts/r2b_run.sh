mkdir -p gpurun_out
(CML_DENSE_TC=1 timeout 900 python -m pytest tests/test_dense_gpu.py -q 2>&1 | tail -5) > gpurun_out/r2b_tests.log
timeout 600 python bench.py --scale 64 --no-sparse-leg --steps 10 --precision 32 > gpurun_out/r2b_cipher32_scale64_tc.json 2>gpurun_out/r2b_err.log
timeout 600 python bench.py --scale 8 --no-sparse-leg --steps 10 --precision 32 > gpurun_out/r2b_cipher32_scale8_tc.json 2>/dev/null
cat gpurun_out/r2b_tests.log; python - <<'PY'
import json
for f in ("cipher32_scale64_tc","cipher32_scale8_tc"):
    try:
        j=json.loads(open(f"gpurun_out/r2b_{f}.json").read().strip().splitlines()[-1])
        print(f, "%.3g"%j["value"], "step %.3f"%j["ms_per_step"], "kernel %.4f"%j["roofline"]["kernel_ms"], "frac %.4f"%j["roofline"]["frac"], j["roofline"]["kernel"][:40], "launches", j["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
PY
