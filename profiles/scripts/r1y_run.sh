mkdir -p gpurun_out
for sc in 8 64; do timeout 600 python bench.py --scale $sc --no-sparse-leg --steps 10 > gpurun_out/r1y_cipher_scale$sc.json 2>/dev/null; done
timeout 600 python bench.py --scale 64 --no-sparse-leg --steps 10 --precision 32 > gpurun_out/r1y_cipher32_scale64.json 2>/dev/null
python - <<'PY'
import json
for f in ("cipher_scale8","cipher_scale64","cipher32_scale64"):
    j=json.loads(open(f"gpurun_out/r1y_{f}.json").read().strip().splitlines()[-1])
    print(f, "%.3g"%j["value"], "step %.3f"%j["ms_per_step"], "kernel %.4f"%j["roofline"]["kernel_ms"], "frac %.4f"%j["roofline"]["frac"], j["totals"]["trellis_arcs"])
PY
