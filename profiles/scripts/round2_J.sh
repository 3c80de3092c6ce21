#!/bin/bash
# round 2 (second session), call J: replicas per hot count slot (the lane kernel's REDs hit 1,088 arc-class slots)
mkdir -p gpurun_out
P=gpurun_out/round2_J
for c in 64 256 1024; do
  CML_HOT_COPIES=$c timeout 300 python bench.py --workload hmm --no-dense --no-sparse-leg --steps 10 > ${P}_lane_c$c.json 2> ${P}_lane_c$c.err
  python -c "
import json; d=json.loads(open('${P}_lane_c$c.json').read().strip().splitlines()[-1]); print('lane copies $c', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity']['ok'])"
done
for c in 64 256; do
  CML_HOT_COPIES=$c timeout 300 python bench.py --legs none --steps 20 > ${P}_cipher_c$c.json 2> ${P}_cipher_c$c.err
  python -c "
import json; d=json.loads(open('${P}_cipher_c$c.json').read().strip().splitlines()[-1]); print('cipher copies $c', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity']['ok'])"
done
CML_HOT_COPIES=256 timeout 300 python bench.py --workload hmm --no-sparse-leg --steps 10 > ${P}_sparse_c256.json 2> ${P}_sparse_c256.err
python -c "
import json; d=json.loads(open('${P}_sparse_c256.json').read().strip().splitlines()[-1]); print('k_fb_sparse copies 256', d['ms_per_step'], d['roofline']['kernel_ms'])"
timeout 300 python bench.py --workload hmm --no-sparse-leg --steps 10 > ${P}_sparse_c64.json 2> ${P}_sparse_c64.err
python -c "
import json; d=json.loads(open('${P}_sparse_c64.json').read().strip().splitlines()[-1]); print('k_fb_sparse copies 64', d['ms_per_step'], d['roofline']['kernel_ms'])"
