mkdir -p gpurun_out
for sc in 2 4 8; do
  CML_DENSE_TC=1 timeout 300 python bench.py --scale $sc --no-sparse-leg --steps 10 --precision 32 > gpurun_out/r2d_tc_s$sc.json 2>/dev/null
  CML_DENSE_TC=0 timeout 300 python bench.py --scale $sc --no-sparse-leg --steps 10 --precision 32 > gpurun_out/r2d_fma_s$sc.json 2>/dev/null
done
python - <<'PY'
import json
for sc in (2,4,8):
    for k in ("tc","fma"):
        j=json.loads(open(f"gpurun_out/r2d_{k}_s{sc}.json").read().strip().splitlines()[-1])
        print(sc*2000, "lines", k, "%.3g"%j["value"], "step %.3f"%j["ms_per_step"], "kernel %.4f"%j["roofline"]["kernel_ms"])
PY
