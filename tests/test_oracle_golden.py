"""The CPU oracle is pinned against the reference's own golden run log:
carmel/carmel-tutorial/commands.trace EM trajectories, composition sizes and trained weights."""
import os
import re

import numpy as np
import pytest

from helpers import golden, run, stage, trajectory_log2


def _check_traj(got, want):
    assert [i for i, _ in got] == [i for i, _ in want]
    for (_, g), (_, w) in zip(got, want):
        # the log prints 6 significant digits
        assert abs(g - w) <= 1e-5 * max(1.0, abs(w)), (g, w)


def test_epron_jpron_trajectory_and_model(oracle_bin, tmp_path):
    fst, data = stage(tmp_path, "epron-jpron.fst", "epron-jpron.data")
    rc, out, err = run(oracle_bin, ["-t", data, fst])
    assert rc == 0, err
    g = golden()["epron_jpron"]
    _check_traj(trajectory_log2(err), g["trajectory_log2"])
    assert "Converged - maximum weight change less than 0.0001 after 5 iterations." in err
    for key, val in g["final_model_excerpt"].items():
        m = re.search(re.escape(key) + r" ([0-9.e+-]+)\)", out)
        assert m, key
        assert abs(float(m.group(1)) - val) <= 1e-9 * val


def test_cipher_cascade_trajectory_and_trained(oracle_bin, tmp_path, golden_dir):
    data, wfsa, fst = stage(tmp_path, "cipher.data", "cipher.wfsa", "cipher.fst")
    rc, out, err = run(oracle_bin, ["--train-cascade", "-HJ", data, wfsa, fst])
    assert rc == 0, err
    g = golden()["cipher"]
    assert f"({g['composed_states']} states / {g['composed_arcs']} arcs)" in err
    _check_traj(trajectory_log2(err), g["trajectory_log2"])
    assert "after 22 iterations" in err

    def weights(path):
        d = {}
        for ln in open(path):
            m = re.match(r'\(0 \(0 ("[^"]*") ("[^"]*") (\S+)\)\)', ln.strip())
            if m:
                w = m.group(3)
                d[(m.group(1), m.group(2))] = float(w[2:]) if w.startswith("e^") else np.log(float(w))
        return d
    got, want = weights(fst + ".trained"), weights(os.path.join(golden_dir, "cipher.fst.trained"))
    assert got.keys() == want.keys() and len(want) > 500
    for k in want:  # learned weights: ln-domain agreement (summation order differs from the 2010 binary)
        assert abs(got[k] - want[k]) <= 1e-6 * max(1.0, abs(want[k])), (k, got[k], want[k])


def test_tagging_cascade_trajectory(oracle_bin, tmp_path):
    data, fsa, fst = stage(tmp_path, "tagging.data", "tagging.fsa", "tagging.fst")
    rc, out, err = run(oracle_bin, ["--train-cascade", "-HJ", data, fsa, fst])
    assert rc == 0, err
    g = golden()["tagging"]
    assert f"({g['composed_states']} states / {g['composed_arcs']} arcs)" in err
    _check_traj(trajectory_log2(err), g["trajectory_log2"])
    assert "after 9 iterations" in err


def test_cluster_first_start(oracle_bin, tmp_path):
    data, fsa = stage(tmp_path, "cluster.data", "cluster.fsa")
    rc, out, err = run(oracle_bin, ["-t", "-M", "2", data, fsa])
    assert rc == 0, err
    _check_traj(trajectory_log2(err)[:2], golden()["cluster_first_start"]["trajectory_log2"])


def test_cat_spellout_cascade_first_start(oracle_bin, tmp_path):
    """`--train-cascade -HJ cluster.data cat.fsa spellout.fst` (golden log `commands.trace:641-649`): the
    composition is 4 states / 316 arcs, the uniform start sits on a saddle, and the cascade needs its third
    iteration before it may stop."""
    data, cat, spell = stage(tmp_path, "cluster.data", "cat.fsa", "spellout.fst")
    rc, out, err = run(oracle_bin, ["--train-cascade", "-HJ", data, cat, spell])
    assert rc == 0, err
    g = golden()["cat_spellout_first_start"]
    assert f"({g['composed_states']} states / {g['composed_arcs']} arcs)" in err
    _check_traj(trajectory_log2(err), g["trajectory_log2"])
    assert "Converged - per-example perplexity ratio exceeds 0.999 after 3 iterations." in err
