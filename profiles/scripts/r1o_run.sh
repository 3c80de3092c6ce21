mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r1o_tests.log
for w in cipher hmm forest gibbs; do timeout 400 python bench.py --workload $w > gpurun_out/r1o_bench_$w.json 2> gpurun_out/r1o_bench_$w.err; done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1o_bench_reference.json 2>&1
cat gpurun_out/r1o_tests.log; for w in cipher hmm forest gibbs; do head -c 300 gpurun_out/r1o_bench_$w.json; echo; done
