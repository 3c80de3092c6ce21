mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r1s_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1s_smoke.log 2>&1
for w in cipher hmm forest gibbs; do timeout 400 python bench.py --workload $w > gpurun_out/r1s_bench_$w.json 2> gpurun_out/r1s_bench_$w.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1s_cipher_launches.csv python bench.py --steps 2 --warmup 3 --no-sparse-leg > gpurun_out/r1s_launches.log 2>&1
cat gpurun_out/r1s_tests.log; tail -2 gpurun_out/r1s_smoke.log; python - <<'PY'
import json
for w in ("cipher","hmm","forest","gibbs"):
    try:
        j=json.loads(open(f"gpurun_out/r1s_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, "%.3g"%j["value"], "step %.3f"%j["ms_per_step"], "kernel %.4f"%j["roofline"]["kernel_ms"], "frac %.4f"%j["roofline"]["frac"], "traffic", j["roofline"]["traffic"])
    except Exception as e:
        print(w, "ERR", e)
PY
