mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12) > gpurun_out/r1p_tests.log
for w in cipher hmm; do timeout 400 python bench.py --workload $w > gpurun_out/r1p_bench_$w.json 2> gpurun_out/r1p_bench_$w.err; done
timeout 400 python bench.py --workload cipher --precision 32 --no-sparse-leg > gpurun_out/r1p_bench_cipher32.json 2> /dev/null
timeout 300 python bench.py --workload gibbs > gpurun_out/r1p_bench_gibbs.json 2> /dev/null
cat gpurun_out/r1p_tests.log; for w in cipher hmm cipher32 gibbs; do head -c 300 gpurun_out/r1p_bench_$w.json; echo; done
