mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_train_gpu.py -q -k "cluster_random_restarts or epron_jpron" 2>&1 | tail -30) > gpurun_out/r2h_tests.log
cat gpurun_out/r2h_tests.log
cd /tmp && cp $GRAFT_REPO_ROOT/tests/golden/cluster.data $GRAFT_REPO_ROOT/tests/golden/cluster.fsa . && (timeout 120 $GRAFT_REPO_ROOT/carmel_b200/_build/carmel-b200 -t -HJ -! 2 -R 3 -M 30 cluster.data cluster.fsa 2>&1 >/dev/null | grep -v "^option" | cut -c1-160 | tail -60) > $GRAFT_REPO_ROOT/gpurun_out/r2h_cluster_restarts.log
