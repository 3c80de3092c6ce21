"""Pins the forest-em CPU oracle (oracle/forest_oracle.hpp).  forest-em ships no expected outputs, so:
 (1) parse -> print round trips of the reference's own unit-test vectors (forest-em/forest.hpp:1041-1070),
 (2) inside scores and expected rule counts against brute-force enumeration of every derivation, on the
     reference's sample forests and on seeded random forests,
 (3) cross-program identity: the cipher cascade exported with --fem-forest/--fem-norm/--fem-param
     (carmel/src/cascade.h:85-166) and trained by the forest oracle must reproduce the likelihood
     trajectory of the reference's golden log (carmel-tutorial/commands.trace:6905-6950)."""
import json
import math
import os

import numpy as np
import pytest

from forest_helpers import (brute_force, parse_ln, random_forest, random_normgroups, read_weights, split_forests)
from helpers import GOLDEN, golden, run, stage

FDIR = os.path.join(GOLDEN, "forest")


def test_parse_print_round_trip(forest_oracle_bin, tmp_path):
    vecs = json.load(open(os.path.join(FDIR, "test_forests.json")))
    src = os.path.join(str(tmp_path), "f")
    open(src, "w").write("\n".join(vecs) + "\n")
    rc, out, err = run(forest_oracle_bin, ["-f", src, "-i", "0", f"--print-forests={tmp_path}/p"])
    assert rc == 0, err
    assert open(f"{tmp_path}/p").read().split("\n")[:-1] == vecs


def _check_against_brute_force(forest_oracle_bin, tmp_path, forests_text, norm_text, ln_w, prec_flag, rel):
    d = str(tmp_path)
    open(f"{d}/f", "w").write(forests_text)
    open(f"{d}/n", "w").write(norm_text)
    open(f"{d}/w", "w").write("".join(f"e^{v:.17g}\n" for v in ln_w))
    rc, out, err = run(forest_oracle_bin, [*prec_flag, "-f", f"{d}/f", "-n", f"{d}/n", "-I", f"{d}/w", "-i", "0", "-S", f"{d}/s"])
    assert rc == 0, err
    rc, out, err = run(forest_oracle_bin, [*prec_flag, "-f", f"{d}/f", "-n", f"{d}/n", "-I", f"{d}/w", "-i", "1", "-O", f"{d}/c"])
    assert rc == 0, err
    w = {i + 1: math.exp(v) for i, v in enumerate(ln_w)}
    inside = [parse_ln(t) for t in open(f"{d}/s").read().split()]
    texts = split_forests(forests_text)
    assert len(inside) == len(texts)
    want_counts = {}
    for t, got in zip(texts, inside):
        total, counts = brute_force(t, w)
        assert abs(got - math.log(total)) <= rel * max(1.0, abs(math.log(total))), (t, got, math.log(total))
        for r, c in counts.items():
            want_counts[r] = want_counts.get(r, 0.0) + c
    got_counts = read_weights(f"{d}/c")
    for r in range(1, len(got_counts) + 1):
        want = want_counts.get(r, 0.0)
        got = math.exp(got_counts[r - 1]) if got_counts[r - 1] > -math.inf else 0.0
        assert abs(got - want) <= rel * 10 * max(1.0, want), (r, got, want)


@pytest.mark.parametrize("prec,rel", [(["-U"], 1e-12), ([], 2e-6)])
def test_sample_forests_brute_force(forest_oracle_bin, tmp_path, prec, rel):
    rng = np.random.default_rng(5)
    forests = open(os.path.join(FDIR, "forests")).read()
    norm = open(os.path.join(FDIR, "norm")).read()
    _check_against_brute_force(forest_oracle_bin, tmp_path, forests, norm, list(np.log(rng.uniform(0.05, 1.0, 16))), prec, rel)


def test_best_forest_brute_force(forest_oracle_bin, tmp_path):
    forests = open(os.path.join(FDIR, "best_forest")).read()
    norm = open(os.path.join(FDIR, "best_norm")).read()
    ln_w = read_weights(os.path.join(FDIR, "best_weights"))
    _check_against_brute_force(forest_oracle_bin, tmp_path, forests, norm, ln_w, ["-U"], 1e-12)


@pytest.mark.parametrize("seed", range(6))
def test_random_forests_brute_force(forest_oracle_bin, tmp_path, seed):
    rng = np.random.default_rng(100 + seed)
    n_rules = 12
    forests = "\n".join(random_forest(rng, n_rules, depth=2) for _ in range(5)) + "\n"
    norm = random_normgroups(rng, n_rules)
    _check_against_brute_force(forest_oracle_bin, tmp_path, forests, norm, list(np.log(rng.uniform(0.05, 1.0, n_rules))),
                               ["-U"], 1e-11)


def test_cipher_cross_program_identity(oracle_bin, forest_oracle_bin, tmp_path):
    data, wfsa, fst = stage(tmp_path, "cipher.data", "cipher.wfsa", "cipher.fst")
    d = str(tmp_path)
    rc, out, err = run(oracle_bin, ["--train-cascade", "--normby=NC", "-HJ", f"--fem-forest={d}/c.forest",
                                    f"--fem-norm={d}/c.norm", f"--fem-param={d}/c.param", data, wfsa, fst])
    assert rc == 0, err
    # forest-em's -e is on the relative change of the average log prob; 1e-3 reproduces carmel's -X .999 stop here
    rc, out, err = run(forest_oracle_bin, ["-U", "-f", f"{d}/c.forest", "-n", f"{d}/c.norm", "-I", f"{d}/c.param",
                                           "-i", "22", "-e", "0", f"--history={d}/h", "-o", f"{d}/c.out"])
    assert rc == 0, err
    want = golden()["cipher"]["trajectory_log2"]
    hist = [ln.split() for ln in open(f"{d}/h")]
    assert len(hist) == len(want) == 22
    n = 10
    for (it, log2p), h in zip(want, hist):
        got = float(h[1]) * n / math.log(2)
        assert int(h[0]) == it and abs(got - log2p) <= 1.01e-5 * abs(log2p), (it, got, log2p)


def test_forest_gibbs_oracle_sanity(forest_oracle_bin, tmp_path):
    """forest-em --crp restatement (forest.hpp:726-816, forest-em.hpp:694-797): PARITY UNPINNED by the reference (no
    expected outputs, boost RNG) -- pinned to first principles instead: (i) every sampled derivation is a derivation of
    its forest, (ii) with a deterministic forest the sample is that derivation and its rules take all the mass,
    (iii) the counts bookkeeping: after the run the weights of every group sum to one."""
    import math
    from forest_helpers import read_weights
    d = str(tmp_path)
    open(f"{d}/f", "w").write("(OR (1 3) (2 3))\n" * 50 + "(4 3)\n" * 10)
    open(f"{d}/n", "w").write("((1 2) (3) (4 5))\n")
    rc, _, err = run(forest_oracle_bin, ["-U", "-f", f"{d}/f", "-n", f"{d}/n", "--crp=40", "--burnin=10", "--seed=3",
                                         f"--outsample-file={d}/s", "-o", f"{d}/w"])
    assert rc == 0, err
    lines = open(f"{d}/s").read().splitlines()
    assert len(lines) == 60
    assert all(ln in ("1 3", "2 3") for ln in lines[:50]) and all(ln == "4 3" for ln in lines[50:])
    w = [math.exp(x) for x in read_weights(f"{d}/w")]
    assert abs(w[0] + w[1] - 1) < 1e-9 and abs(w[2] - 1) < 1e-9 and abs(w[3] + w[4] - 1) < 1e-9
    assert w[3] > 0.95  # rule 4 is used by 10 forests, rule 5 by none: alpha * p0 * N = 0.1 pseudo-counts against 10
    assert err.count("Gibbs i=") == 41
