#!/bin/bash
# round 2 (second session), call I: forest iteration launch list (what the 0.2 ms around the kernel is), maxdiff fix
mkdir -p gpurun_out
P=gpurun_out/round2_I
timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -x -q -k "forest_tiles or cyclic" > ${P}_tests.log 2>&1
echo "tests rc=$?"; tail -3 ${P}_tests.log
CB200_NO_CPU=1 timeout 300 python bench.py --workload forest --steps 10 > ${P}_default.json 2> ${P}_default.err
python -c "
import json; d=json.loads(open('${P}_default.json').read().strip().splitlines()[-1]); print('forest', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
CB200_NO_CPU=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_forest_launches.csv \
  python bench.py --workload forest --steps 2 --warmup 3 > ${P}_launches_bench.log 2>&1
echo "launch list rc=$?"
