mkdir -p gpurun_out
export CML_SPARSE_MIN_SEQ=1
mc() { echo "=== memcheck(all processes) $*" ; timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 99 --print-limit 5 "$@" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|error" | sort | uniq -c | tail -8; }
rc() { echo "=== racecheck $*" ; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 5 "$@" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|hazard" | sort | uniq -c | tail -8; }
(
mc python -m pytest tests/test_dense_gpu.py -q -x -k "synthetic_27_state and mode0"
mc python -m pytest tests/test_forest_gpu.py -q -x -k "zero_probability_forests and mode0"
mc python -m pytest tests/test_sparse_gpu.py -q -x -k "without_final_weights"
python - <<'PY'
import os, tempfile
from carmel_b200 import synth
d = tempfile.mkdtemp()
w = synth.write_cipher(d, n_lines=40, line_len=12, seed=3)
open("/tmp/gd_files.txt", "w").write(" ".join(w["files"]))
PY
mc carmel_b200/_build/carmel-b200 --crp -M 3 --crp-batched --priors=0,1e-2 --seed=2 -q $(cat /tmp/gd_files.txt)
rc python -m pytest tests/test_dense_gpu.py -q -x -k "estep_equals"
rc python -m pytest tests/test_sparse_gpu.py -q -x -k "estep_equals"
rc python -m pytest tests/test_lane_gpu.py -q -x -k "layout_is_used"
) > gpurun_out/r1v_sanitizer.log 2>&1
cat gpurun_out/r1v_sanitizer.log
