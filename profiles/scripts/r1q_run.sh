mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_estep_gpu.py tests/test_train_gpu.py -q 2>&1 | tail -8) > gpurun_out/r1q_tests.log
timeout 400 python bench.py --workload cipher > gpurun_out/r1q_bench_cipher.json 2> gpurun_out/r1q_bench_cipher.err
(timeout 900 python -m pytest tests/test_forest_gpu.py -q 2>&1 | tail -8) > gpurun_out/r1q_forest_tests.log
timeout 400 python bench.py --workload forest > gpurun_out/r1q_bench_forest.json 2> gpurun_out/r1q_bench_forest.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_forest_thread --launch-skip 3 --launch-count 1 -f -o gpurun_out/r1q_forest python bench.py --workload forest --steps 1 --warmup 3 > gpurun_out/r1q_ncu_forest.log 2>&1
cat gpurun_out/r1q_tests.log gpurun_out/r1q_forest_tests.log; python - <<'PY'
import json
j=json.loads(open("gpurun_out/r1q_bench_cipher.json").read().strip().splitlines()[-1])
print("dense", j["value"], j["ms_per_step"], j["roofline"]["kernel_ms"]); sp=j["sparse_path"]; print("sparse", sp["value"], sp["ms_per_step"], sp["roofline"]["kernel_ms"], sp["roofline"]["frac"])
j=json.loads(open("gpurun_out/r1q_bench_forest.json").read().strip().splitlines()[-1])
print("forest", j["value"], j["ms_per_step"], j["roofline"]["kernel_ms"], j["roofline"]["frac"])
PY
