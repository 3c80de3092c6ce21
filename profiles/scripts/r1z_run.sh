mkdir -p gpurun_out
(CML_DENSE_TC=1 timeout 900 python -m pytest tests/test_dense_gpu.py -q 2>&1 | tail -15) > gpurun_out/r1z_tests.log
timeout 600 python bench.py --scale 64 --no-sparse-leg --steps 10 --precision 32 > gpurun_out/r1z_cipher32_scale64_tc.json 2>gpurun_out/r1z_err.log
CML_DENSE_TC=0 timeout 600 python bench.py --scale 64 --no-sparse-leg --steps 10 --precision 32 > gpurun_out/r1z_cipher32_scale64_fma.json 2>/dev/null
CML_DENSE_TC=1 timeout 300 python bench.py --no-sparse-leg --steps 10 --precision 32 > gpurun_out/r1z_cipher32_scale1_tc.json 2>/dev/null
cat gpurun_out/r1z_tests.log; python - <<'PY'
import json
for f in ("cipher32_scale64_tc","cipher32_scale64_fma","cipher32_scale1_tc"):
    try:
        j=json.loads(open(f"gpurun_out/r1z_{f}.json").read().strip().splitlines()[-1])
        print(f, "%.3g"%j["value"], "step %.3f"%j["ms_per_step"], "kernel %.4f"%j["roofline"]["kernel_ms"], "launches", j["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r1z_err.log
