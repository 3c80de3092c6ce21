mkdir -p gpurun_out
timeout 170 python bench.py > gpurun_out/r2l_bench_cipher.json 2> gpurun_out/r2l_bench_cipher.err
tail -c 2500 gpurun_out/r2l_bench_cipher.json
