#!/bin/bash
# round 2 (second session), call P: the Gibbs leg with its parity entry (sequential-mode derivations vs the oracle)
mkdir -p gpurun_out
timeout 110 python bench.py --workload gibbs --steps 3 > gpurun_out/round2_P_gibbs.json 2> gpurun_out/round2_P_gibbs.err; echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/round2_P_gibbs.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['parity'])"
