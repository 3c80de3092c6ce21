"""N>1 data path on CPU: world_size-2 (and 3) torch.distributed `gloo` jobs over the product's shard partitioner.

Identity under test: examples / forests are independent given the weights, so the all-reduced
[counts | sum ln p | n] of the shards must equal the unsharded E-step (here: the CPU oracle on the whole corpus),
and the shards must partition the corpus (cover, no overlap, order preserved)."""
import json
import math
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from forest_helpers import read_weights
from helpers import GOLDEN, ROOT, read_estimate_dump, run, stage

WORKER = os.path.join(ROOT, "tests", "gloo_worker.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(cfg_path, world):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, WORKER, cfg_path], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                      text=True))
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err[-3000:]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["epron", "cipher"])
def test_carmel_shards_allreduce_to_unsharded(native_lib, oracle_bin, tmp_path, case, world):
    from carmel_b200 import CLI_PATH
    d = str(tmp_path)
    if case == "epron":
        fst, data = stage(tmp_path, "epron-jpron.fst", "epron-jpron.data")
        args, files = ["-t"], [data, fst]
    else:
        data, wfsa, fst = stage(tmp_path, "cipher.data", "cipher.wfsa", "cipher.fst")
        # two lines of the ciphertext keep the pure-python forward-backward of the worker fast
        lines = open(data).read().split("\n")
        open(data, "w").write("\n".join(lines[:10]) + "\n")  # 5 whole pairs (the LM only accepts whole sentences)
        args, files = ["--train-cascade"], [data, wfsa, fst]
    rc, out, err = run(oracle_bin, [*args, f"--dump-estimate={d}/est", *files])
    assert rc == 0, err
    est = read_estimate_dump(f"{d}/est")
    cfg = {"case": "carmel", "dir": d, "cli": CLI_PATH, "args": args, "files": files, "ln_w": est["ln_w"].tolist()}
    json.dump(cfg, open(f"{d}/cfg.json", "w"))
    _launch(f"{d}/cfg.json", world)
    res = json.load(open(f"{d}/result.json"))
    red = np.asarray(res["reduced"])
    assert sum(res["per_rank"]) == len(est["ln_p"]) == red[-1]
    assert sum(1 for n in res["per_rank"] if n > 0) >= 2          # really sharded
    assert abs(red[-2] - est["ln_p"].sum()) <= 1e-9 * abs(est["ln_p"].sum())
    want = np.exp(est["ln_counts"])
    np.testing.assert_allclose(red[:-2], want, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("world", [2, 3])
def test_forest_shards_allreduce_to_unsharded(native_lib, forest_oracle_bin, tmp_path, world):
    from carmel_b200 import FOREST_CLI_PATH
    d = str(tmp_path)
    forests, norm = os.path.join(GOLDEN, "forest", "forests"), os.path.join(GOLDEN, "forest", "norm")
    rng = np.random.default_rng(3)
    ln_w = [-math.inf] + list(np.log(rng.uniform(0.05, 1.0, 16)))
    open(f"{d}/w", "w").write("".join(f"e^{v:.17g}\n" for v in ln_w[1:]))
    rc, out, err = run(forest_oracle_bin, ["-U", "-f", forests, "-n", norm, "-I", f"{d}/w", "-i", "1", "-O", f"{d}/c", f"--history={d}/h"])
    assert rc == 0, err
    cfg = {"case": "forest", "dir": d, "cli": FOREST_CLI_PATH, "forests": forests, "norm": norm, "ln_w": ln_w}
    json.dump(cfg, open(f"{d}/cfg.json", "w"))
    _launch(f"{d}/cfg.json", world)
    res = json.load(open(f"{d}/result.json"))
    red = np.asarray(res["reduced"])
    assert sum(res["per_rank"]) == 5 == red[-1] and all(n > 0 for n in res["per_rank"])
    h = open(f"{d}/h").read().split()
    assert abs(red[-2] / red[-1] - float(h[1])) <= 1e-12 * abs(float(h[1]))
    want = np.exp(np.asarray(read_weights(f"{d}/c")))
    np.testing.assert_allclose(red[1:-2], want, rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("world", [2, 3])
def test_batched_gibbs_count_deltas_allreduce_to_unsharded(tmp_path, world):
    """sharded batched --crp sweeps: the all-reduced per-rank delta tables, added to the replicated counts,
    must equal the one-GPU update (remove every block's old sample, add its new one)"""
    rng = np.random.default_rng(17)
    n_params, n_norms, n_arcs = 40, 6, 25
    norm = [int(rng.integers(0, n_norms)) if rng.random() < 0.8 else -1 for _ in range(n_params)]
    chains = [[int(p) for p in rng.choice(n_params, size=int(rng.integers(1, 4)), replace=False)] for _ in range(n_arcs)]
    blocks = [([int(a) for a in rng.integers(0, n_arcs, int(rng.integers(0, 9)))],
               [int(a) for a in rng.integers(0, n_arcs, int(rng.integers(1, 9)))], float(rng.choice([1.0, 2.0, 0.5])))
              for _ in range(23)]
    d = str(tmp_path)
    json.dump({"case": "gibbs", "dir": d, "n_params": n_params, "n_norms": n_norms, "param_norm": norm, "chains": chains,
               "blocks": blocks}, open(f"{d}/cfg.json", "w"))
    _launch(f"{d}/cfg.json", world)
    res = json.load(open(f"{d}/result.json"))
    red = np.asarray(res["reduced"])
    assert sum(res["per_rank"]) == len(blocks) == red[-1] and all(n > 0 for n in res["per_rank"])
    count, normsum = np.zeros(n_params), np.zeros(n_norms)
    for old, new, wt in blocks:  # the sequentially applied update of one GPU (k_gibbs_apply)
        for arcs, sign in ((old, -wt), (new, wt)):
            for a in arcs:
                for p in chains[a]:
                    if norm[p] >= 0:
                        count[p] += sign
                        normsum[norm[p]] += sign
    np.testing.assert_allclose(red[:n_params], count, rtol=0, atol=1e-12)
    np.testing.assert_allclose(red[n_params:n_params + n_norms], normsum, rtol=0, atol=1e-12)


def test_shard_partition_covers_corpus(native_lib, tmp_path):
    """blocks of --shard=r/N concatenate to the unsharded lattice dump, for several N"""
    from carmel_b200 import CLI_PATH
    data, fsa, fst = stage(tmp_path, "tagging.data", "tagging.fsa", "tagging.fst")
    lines = open(data).read().split("\n")[:120]
    open(data, "w").write("\n".join(lines) + "\n")
    d = str(tmp_path)
    rc, out, err = run(CLI_PATH, ["--train-cascade", "--trellis-only", f"--dump-trellis={d}/all", data, fsa, fst])
    assert rc == 0, err
    whole = open(f"{d}/all", "rb").read()
    for n in (2, 4, 7):
        parts, count = [], 0
        for r in range(n):
            rc, out, err = run(CLI_PATH, ["--train-cascade", "--trellis-only", f"--shard={r}/{n}", f"--dump-trellis={d}/p{r}", data, fsa, fst])
            assert rc == 0, err
            b = open(f"{d}/p{r}", "rb").read()
            count += int.from_bytes(b[:4], "little")
            parts.append(b[8:])
        assert count == int.from_bytes(whole[:4], "little")
        assert b"".join(parts) == whole[8:]
