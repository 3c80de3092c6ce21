#!/bin/bash
# round 2, call K: the restructured bench (all legs) on one GPU + ABI/GPU tests of the new entry points
mkdir -p gpurun_out
START=$(date +%s)
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
echo "bench rc=$? wall $(( $(date +%s) - START ))s"
tail -5 gpurun_out/r2k_bench.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r2k_bench.json"))
def show(name, d):
    if not isinstance(d, dict): print(name, d); return
    if "failed" in d: print(name, "FAILED", d["failed"]); return
    r=d.get("roofline") or {}
    print(name, "value=%.4g" % d.get("value", float("nan")) if d.get("value") else "", "ms/step", d.get("ms_per_step"), "frac", r.get("frac"), "kernel_ms", r.get("kernel_ms"), "parity", d.get("parity"), "wall", d.get("wall_s"))
show("main", j); print("e2e", j["e2e"]["value"], j["e2e"]["ms_per_step"]); print("launches", j["gpu_launches"], "parity_ok", j.get("parity_ok"), j.get("parity_failures"))
for k in ("dense_path","c3","c5","c4"): show(k, j.get(k))
if isinstance(j.get("c3"), dict): show("c3.dense", j["c3"].get("dense_path"))
print("cli", j.get("e2e_cli")); print("cpu", j.get("cpu_baseline")); print("tf32", j.get("tf32_peak_tflops_measured"))
PY
