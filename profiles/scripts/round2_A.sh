#!/bin/bash
# round 2 (second session), call A: ncu launch list of the default bench command + ncu --set full captures of the
# kernels as shipped (k_fb_wide on the cipher, k_fb_lane / k_fb_sparse on hmm, k_forest_thread on the all-distinct corpus)
mkdir -p gpurun_out
P=gpurun_out/round2_A
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${P}_launches.csv \
  python bench.py --steps 2 --warmup 3 --legs none > ${P}_launches_bench.log 2>&1
echo "launch list rc=$?"
cap() {  # name, kernel regex, skip, bench args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip -c 1 -f -o ${P}_$name \
    python bench.py "$@" > ${P}_$name.log 2>&1
  echo "$name rc=$?"
}
cap k_fb_wide k_fb_wide 4 --steps 2 --legs none
cap k_forest_thread k_forest_thread 3 --workload forest --steps 2
cap k_fb_lane k_fb_lane 3 --workload hmm --no-dense --no-sparse-leg --steps 2
cap k_fb_sparse k_fb_sparse 3 --workload hmm --no-sparse-leg --steps 2
ls -la gpurun_out | tail -12
