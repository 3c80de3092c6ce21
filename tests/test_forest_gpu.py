"""forest-em parity on the GPU through the product command line
forest-em-b200 (host C++ -> C ABI cml_forests_* -> CUDA) against the CPU oracle and the golden log.

Tolerances (north_star): average log-likelihood per iteration and learned weights within 1e-6 relative with
-U (fp64 scores), 1e-4 in the default fp32 mode (forest-em's own default precision)."""
import math
import os

import numpy as np
import pytest

from forest_helpers import parse_ln, random_forest, random_normgroups, read_weights
from helpers import GOLDEN, golden, run, stage

pytestmark = pytest.mark.gpu
FDIR = os.path.join(GOLDEN, "forest")
MODES = [(["-U"], 1e-6), ([], 1e-4)]


@pytest.fixture(scope="module")
def fem(native_lib):
    from carmel_b200 import FOREST_CLI_PATH
    return FOREST_CLI_PATH


def _hist(path):
    return [(int(r[0]), float(r[1]), float(r[2]), int(r[3]), int(r[4])) for r in (ln.split() for ln in open(path)) if r]


def _close_ln(a, b, rel, floor=-60.0):
    if a == b:
        return True
    if a < floor and b < floor:  # both negligible probabilities / counts
        return True
    return abs(math.expm1(a - b)) <= rel if abs(a - b) < 1 else False


def _compare(fem, forest_oracle_bin, tmp_path, args, mode, rel, check_index=True, layouts=("group", "thread", "level")):
    """product (every device layout family) against the oracle: history, weights, counts, per-forest inside"""
    d = str(tmp_path)
    rc, out, oerr = run(forest_oracle_bin, [*mode, *args, "-o", f"{d}/o.w", "-O", f"{d}/o.c", "-S", f"{d}/o.s", f"--history={d}/o.h"])
    assert rc == 0, oerr
    ho = _hist(f"{d}/o.h")
    err = ""
    for layout in layouts:
        rc, out, err = run(fem, [*mode, *args, f"--layout={layout}", "-o", f"{d}/p.w", "-O", f"{d}/p.c", "-S", f"{d}/p.s",
                                 f"--history={d}/p.h"])
        assert rc == 0, err
        hp = _hist(f"{d}/p.h")
        assert len(hp) == len(ho), (layout, len(hp), len(ho), err[-2000:])
        for a, b in zip(hp, ho):
            assert a[0] == b[0] and a[4] == b[4], (layout, a, b)
            assert abs(a[1] - b[1]) <= rel * max(1.0, abs(b[1])), (layout, a, b)
            assert abs(a[2] - b[2]) <= 20 * rel * max(1.0, abs(b[2])) + 1e-12, (layout, a, b)
            if check_index and rel <= 1e-6 and b[2] > 1e-3:
                assert a[3] == b[3], (layout, a, b)
        # per-forest ln inside: in fp32 both sides carry the rounding of intermediate logs that are much larger in
        # magnitude than the result (sums over thousands of derivations), so the float-vs-float bound is looser
        for name, tol in (("w", 20 * rel), ("c", 20 * rel), ("s", rel if rel <= 1e-6 else 10 * rel)):
            got, want = read_weights(f"{d}/p.{name}"), read_weights(f"{d}/o.{name}")
            assert len(got) == len(want), (layout, name, len(got), len(want))
            for i, (a, b) in enumerate(zip(got, want)):
                if name == "s":
                    assert (a == b) or abs(a - b) <= tol * max(1.0, abs(b)), (layout, name, i, a, b)
                else:
                    assert _close_ln(a, b, tol), (layout, name, i + 1, a, b)
    return err


@pytest.mark.parametrize("mode,rel", MODES)
def test_sample_forests(fem, forest_oracle_bin, tmp_path, mode, rel):
    args = ["-f", os.path.join(FDIR, "forests"), "-n", os.path.join(FDIR, "norm"), "-i", "12"]
    err = _compare(fem, forest_oracle_bin, tmp_path, args, mode, rel)
    assert "i=12: probability=2^" in err


@pytest.mark.parametrize("mode,rel", MODES)
def test_norm_and_forests_one_stream(fem, forest_oracle_bin, tmp_path, mode, rel):
    both = os.path.join(FDIR, "norm_and_forests")
    _compare(fem, forest_oracle_bin, tmp_path, ["-f", both, "-n", both, "-i", "8", "-p", "0.1", "-k", "0.5"], mode, rel)


def test_best_forest_initparams(fem, forest_oracle_bin, tmp_path):
    args = ["-f", os.path.join(FDIR, "best_forest"), "-n", os.path.join(FDIR, "best_norm"), "-I", os.path.join(FDIR, "best_weights"),
            "-i", "6", "-N"]
    _compare(fem, forest_oracle_bin, tmp_path, args, ["-U"], 1e-6)


def _random_corpus(tmp_path, seed, n_forests, n_rules, depth):
    rng = np.random.default_rng(seed)
    d = str(tmp_path)
    open(f"{d}/f", "w").write("\n".join(random_forest(rng, n_rules, depth=depth) for _ in range(n_forests)) + "\n")
    open(f"{d}/n", "w").write(random_normgroups(rng, n_rules))
    return f"{d}/f", f"{d}/n"


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("mode,rel", MODES)
def test_random_forests(fem, forest_oracle_bin, tmp_path, seed, mode, rel):
    f, n = _random_corpus(tmp_path, 300 + seed, n_forests=120, n_rules=60, depth=5)
    extra = [[], ["-z"], ["-p", "0.01"], ["-u"]][seed]
    # rules outside every normalization group keep weight 0 unless -u: most forests then have zero probability,
    # which exercises the "0 prob removed" path
    _compare(fem, forest_oracle_bin, tmp_path, ["-f", f, "-n", n, "-i", "6", *extra], mode, rel, check_index=False)


@pytest.mark.parametrize("mode,rel", MODES)
def test_random_forests_thread_layout_without_stacks(fem, forest_oracle_bin, tmp_path, monkeypatch, mode, rel):
    """the thread-per-forest kernel reads tree children / tree parents from per-lane shared-memory stacks; with
    CML_FOREST_NO_STACK every value comes from the global arrays (the path forests in non-tree text order take)"""
    monkeypatch.setenv("CML_FOREST_NO_STACK", "1")
    f, n = _random_corpus(tmp_path, 311, n_forests=150, n_rules=50, depth=6)
    _compare(fem, forest_oracle_bin, tmp_path, ["-f", f, "-n", n, "-i", "5", "-u"], mode, rel, check_index=False,
             layouts=("thread",))


@pytest.mark.parametrize("mode,rel", MODES)
@pytest.mark.parametrize("small_kb", ["0", "1"])
def test_level_layout_large_tiles(fem, forest_oracle_bin, tmp_path, monkeypatch, mode, rel, small_kb):
    """the level layout has two tile classes (128-thread CTAs for runs of forests that fit 12 KB of values, 512-thread
    CTAs for larger forests); with the small class switched off (0) or shrunk to 1 KB the same corpus runs on the large
    tiles, alone or mixed with small ones"""
    monkeypatch.setenv("CML_FOREST_LEVEL_SMEM_KB", small_kb)
    rng = np.random.default_rng(913)
    d = str(tmp_path)
    n_rules = 80
    open(f"{d}/f", "w").write("\n".join(random_forest(rng, n_rules, depth=int(rng.integers(2, 9)), share=0.3) for _ in range(96)) + "\n")
    open(f"{d}/n", "w").write(random_normgroups(rng, n_rules))
    _compare(fem, forest_oracle_bin, tmp_path, ["-f", f"{d}/f", "-n", f"{d}/n", "-i", "4", "-u"], mode, rel,
             check_index=False, layouts=("level",))


@pytest.mark.parametrize("mode,rel", MODES)
def test_deep_random_forests_thread_layout(fem, forest_oracle_bin, tmp_path, mode, rel):
    """deeper forests with many shared sub-forests: value stack and path stack several levels deep, back-reference
    links mixed with stack links"""
    rng = np.random.default_rng(912)
    d = str(tmp_path)
    n_rules = 80
    open(f"{d}/f", "w").write("\n".join(random_forest(rng, n_rules, depth=8, share=0.3) for _ in range(64)) + "\n")
    open(f"{d}/n", "w").write(random_normgroups(rng, n_rules))
    _compare(fem, forest_oracle_bin, tmp_path, ["-f", f"{d}/f", "-n", f"{d}/n", "-i", "4", "-u"], mode, rel,
             check_index=False, layouts=("thread",))


@pytest.mark.parametrize("mode,rel", MODES)
def test_zero_probability_forests(fem, forest_oracle_bin, tmp_path, mode, rel):
    d = str(tmp_path)
    # rule 5 is in no normalization group => weight 0 => forests 3 and 5 have no derivation with non-zero probability
    open(f"{d}/f", "w").write("(1 2 3)\n(OR 3 (4 1))\n(5 5)\n(OR (1 #1(OR 3 4)) (2 #1))\n(OR (5 1) (2 5))\n")
    open(f"{d}/n", "w").write("((1 2) (3 4))\n")
    # (two-member groups change by equal and opposite amounts: which member holds the "largest" change is a rounding tie)
    err = _compare(fem, forest_oracle_bin, tmp_path, ["-f", f"{d}/f", "-n", f"{d}/n", "-i", "5"], mode, rel, check_index=False)
    assert "N=3 (2 0 prob removed)" in err
    assert "Warning: 0 probability for forest #3" in err and "Warning: 0 probability for forest #5" in err


def test_big_forests_cta_class(fem, forest_oracle_bin, tmp_path):
    # deep forests exceed the shared-memory classes and take the one-CTA-per-forest kernel
    rng = np.random.default_rng(77)
    d = str(tmp_path)
    n_rules = 200
    open(f"{d}/f", "w").write("\n".join(random_forest(rng, n_rules, depth=9, share=0.05) for _ in range(6)) + "\n")
    ids = list(range(1, n_rules + 1))
    open(f"{d}/n", "w").write("(" + " ".join("(" + " ".join(map(str, ids[i:i + 8])) + ")" for i in range(0, n_rules, 8)) + ")\n")
    err = _compare(fem, forest_oracle_bin, tmp_path, ["-f", f"{d}/f", "-n", f"{d}/n", "-i", "4"], ["-U"], 1e-6, check_index=False)
    assert "N=6" in err


def test_cipher_forests_golden_trajectory(fem, oracle_bin, tmp_path):
    """carmel's cipher cascade exported as forests and trained on the GPU reproduces the reference's golden log"""
    data, wfsa, fst = stage(tmp_path, "cipher.data", "cipher.wfsa", "cipher.fst")
    d = str(tmp_path)
    rc, out, err = run(oracle_bin, ["--train-cascade", "--normby=NC", "-HJ", f"--fem-forest={d}/c.forest", f"--fem-norm={d}/c.norm",
                                    f"--fem-param={d}/c.param", data, wfsa, fst])
    assert rc == 0, err
    want = golden()["cipher"]["trajectory_log2"]
    for layout in ("group", "thread", "level"):
        rc, out, err = run(fem, ["-U", "-f", f"{d}/c.forest", "-n", f"{d}/c.norm", "-I", f"{d}/c.param", "-i", "22", "-e", "0",
                                 f"--layout={layout}", f"--history={d}/h"])
        assert rc == 0, err
        hist = _hist(f"{d}/h")
        assert len(hist) == len(want) == 22
        for (it, log2p), h in zip(want, hist):
            got = h[1] * 10 / math.log(2)
            assert h[0] == it and abs(got - log2p) <= 1.01e-5 * abs(log2p), (layout, it, got, log2p)


def test_errors(fem, tmp_path):
    d = str(tmp_path)
    open(f"{d}/bad", "w").write("(OR (1 2) (3 #9))\n")
    open(f"{d}/n", "w").write("((1 2 3))\n")
    rc, out, err = run(fem, ["-f", f"{d}/bad", "-n", f"{d}/n", "-i", "2"])
    assert rc == 1 and "undefined #9" in err
    open(f"{d}/cyc", "w").write("#1(1 #1 (2 #1 (1 1)))\n")
    rc, out, err = run(fem, ["-f", f"{d}/cyc", "-n", f"{d}/n", "-i", "2"])
    assert rc == 1 and "cyclic" in err
    rc, out, err = run(fem, ["-n", f"{d}/n", "-i", "2"])
    assert rc == 1 and "Missing forests-file" in err


def test_random_restarts(fem, tmp_path):
    """forest-em -r n (forest-em-params.hpp:103, em.hpp:114-214, forest-em.hpp:393-399,660-672): the first start
    is the run without -r; every further start draws each normalisation group's members uniformly on (0,1] and
    normalises them; the parameters written are those of the best iteration of any start.  The draws differ from
    the reference's generator, so the restarts are checked through these properties."""
    d = str(tmp_path)
    rng = np.random.default_rng(77)
    f, n = f"{d}/f", f"{d}/n"
    open(f, "w").write("\n".join(random_forest(rng, 40, depth=5) for _ in range(200)) + "\n")
    open(n, "w").write(random_normgroups(rng, 40, leave_out=0))  # every rule is in a group: all starts are normalised
    args = ["-U", "-f", f, "-n", n, "-i", "10"]
    rc, _, err0 = run(fem, [*args, "-o", f"{d}/w0", f"--history={d}/h0"])
    assert rc == 0, err0
    rc, _, err = run(fem, [*args, "-r", "3", "-s", "11", "-o", f"{d}/w", f"--history={d}/h"])
    assert rc == 0, err
    h0, h = _hist(f"{d}/h0"), _hist(f"{d}/h")
    starts = [k for k, row in enumerate(h) if row[0] == 1] + [len(h)]
    assert len(starts) == 5
    for n_left in (2, 1, 0):
        assert f"Random restart - {n_left} remaining." in err
    assert len(h0) == starts[1]
    for a, b in zip(h0, h[:starts[1]]):
        assert a[0] == b[0] and abs(a[1] - b[1]) <= 1e-12 * abs(a[1])
    firsts = set()
    for a, b in zip(starts, starts[1:]):
        alps = [row[1] for row in h[a:b]]
        assert all(y >= x - 1e-9 * abs(x) for x, y in zip(alps, alps[1:]))  # EM never lowers the likelihood
        firsts.add(round(alps[0], 9))
    assert len(firsts) == 4  # four different starting points
    best = max(row[1] for row in h)
    # the written parameters are the best iteration's: one more estimate from them gives that likelihood
    rc, _, err1 = run(fem, ["-U", "-f", f, "-n", n, "-I", f"{d}/w", "-i", "1", f"--history={d}/h1"])
    assert rc == 0, err1
    assert abs(_hist(f"{d}/h1")[0][1] - best) <= 1e-8 * abs(best)
    # same seed: same draws; another seed: other draws
    rc, _, err2 = run(fem, [*args, "-r", "3", "-s", "11", f"--history={d}/h2"])
    assert rc == 0, err2
    h2 = _hist(f"{d}/h2")
    assert len(h2) == len(h) and all(abs(a[1] - b[1]) <= 1e-9 * abs(a[1]) for a, b in zip(h, h2))
    rc, _, err3 = run(fem, [*args, "-r", "1", "-s", "12", f"--history={d}/h3"])
    assert rc == 0, err3
    assert abs(_hist(f"{d}/h3")[starts[1]][1] - h[starts[1]][1]) > 1e-9


def test_checkpoints_and_resume(fem, tmp_path):
    """forest-em's checkpoint files (forest-em.hpp:166-201: <prefix>.params/.counts.restart.R.iteration.I on every watch
    iteration, -c -x prefix -W period) and resume by -I: 6 iterations in one run == 3 iterations, then 3 more started
    from the iteration-3 parameter checkpoint."""
    rng = np.random.default_rng(20261717)
    d = str(tmp_path)
    forests = [random_forest(rng, n_rules=30, depth=4) for _ in range(60)]
    open(f"{d}/f", "w").write("\n".join(forests) + "\n")
    open(f"{d}/n", "w").write(random_normgroups(rng, 30))
    base = ["-U", "-f", f"{d}/f", "-n", f"{d}/n", "-e", "0"]
    rc, _, err = run(fem, [*base, "-i", "6", "-c", "-x", f"{d}/ck", "-W", "2", "-o", f"{d}/w6", f"--history={d}/h6"])
    assert rc == 0, err
    have = sorted(f for f in os.listdir(d) if f.startswith("ck."))
    # iterations 0,1,2 (<= watch period) and 4 are watch iterations: files are numbered iteration+1
    want_its = [1, 2, 3, 5]
    assert have == sorted([f"ck.{k}.restart.1.iteration.{i}" for k in ("params", "counts") for i in want_its]), have
    rc, _, err = run(fem, [*base, "-i", "3", "-I", f"{d}/ck.params.restart.1.iteration.3", "-o", f"{d}/w33", f"--history={d}/h33"])
    assert rc == 0, err
    h6, h33 = _hist(f"{d}/h6"), _hist(f"{d}/h33")
    assert len(h6) == 6 and len(h33) == 3
    for a, b in zip(h6[3:], h33):
        assert abs(a[1] - b[1]) <= 1e-9 * max(1.0, abs(a[1])), (a, b)
    got, want = read_weights(f"{d}/w33"), read_weights(f"{d}/w6")
    assert len(got) == len(want)
    for x, y in zip(got, want):
        assert _close_ln(x, y, 1e-8), (x, y)


@pytest.mark.parametrize("mode,rel", MODES)
def test_viterbi_matches_oracle(fem, forest_oracle_bin, tmp_path, mode, rel):
    """final viterbi decoding (-v; forest.hpp:507-631, forest-em.hpp:535-550): 'best/sum=pct% tree' per forest"""
    rng = np.random.default_rng(20261818)
    d = str(tmp_path)
    forests = [random_forest(rng, n_rules=40, depth=5, share=0.2) for _ in range(80)]
    open(f"{d}/f", "w").write("\n".join(forests) + "\n")
    open(f"{d}/n", "w").write(random_normgroups(rng, 40))
    args = [*mode, "-f", f"{d}/f", "-n", f"{d}/n", "-i", "4"]
    rc, _, err = run(forest_oracle_bin, [*args, "-v", f"{d}/o.v"])
    assert rc == 0, err
    rc, _, err = run(fem, [*args, "-v", f"{d}/p.v"])
    assert rc == 0, err
    lo, lp = open(f"{d}/o.v").read().splitlines(), open(f"{d}/p.v").read().splitlines()
    assert len(lo) == len(lp) == 80
    same_tree = 0
    for a, b in zip(lp, lo):
        ha, ta = a.split("% ", 1)
        hb, tb = b.split("% ", 1)
        best_a, sum_a = (parse_ln(t) for t in ha.split("=")[0].split("/"))
        best_b, sum_b = (parse_ln(t) for t in hb.split("=")[0].split("/"))
        assert _close_ln(best_a, best_b, 10 * rel, floor=-600.0), (a, b)
        assert _close_ln(sum_a, sum_b, 10 * rel, floor=-600.0), (a, b)
        same_tree += ta == tb
    # in fp64 the derivations are the same; in fp32 the oracle's float scores can tie-break an OR node differently
    assert same_tree == 80 if rel <= 1e-6 else same_tree >= 76, same_tree


def test_forest_cli_two_gpus_equals_one(fem, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(20261919)
    d = str(tmp_path)
    open(f"{d}/f", "w").write("\n".join(random_forest(rng, n_rules=50, depth=5) for _ in range(301)) + "\n")
    open(f"{d}/n", "w").write(random_normgroups(rng, 50))
    base = ["-U", "-f", f"{d}/f", "-n", f"{d}/n", "-i", "5", "-e", "0"]
    rc, _, err = run(fem, [*base, "-o", f"{d}/w1", f"--history={d}/h1"])
    assert rc == 0, err
    rc, _, err = run(fem, [*base, "--gpus=2", "-o", f"{d}/w2", f"--history={d}/h2"], timeout=300)
    assert rc == 0, err
    h1, h2 = _hist(f"{d}/h1"), _hist(f"{d}/h2")
    assert len(h1) == len(h2) == 5
    for a, b in zip(h1, h2):
        assert abs(a[1] - b[1]) <= 1e-9 * max(1.0, abs(a[1])), (a, b)
    for x, y in zip(read_weights(f"{d}/w2"), read_weights(f"{d}/w1")):
        assert _close_ln(x, y, 1e-8), (x, y)


# ---- forest-em --crp (row a24): Gibbs sampling over forests ---------------------------------------------------------
GIBBS_VARIANTS = [
    ["--crp=6"],
    ["--crp=8", "--burnin=3"],
    ["--crp=5", "--final-counts"],
    ["--crp=6", "--burnin=2", "--crp-exclude-prior"],
    ["--crp=5", "--uniform-p0", "--const-alpha=0.5"],
    ["--crp=6", "--high-temp=3", "--low-temp=0.5"],
    ["--crp=4", "--sample-prob", "--const-alpha=2"],
]


@pytest.mark.parametrize("variant", range(len(GIBBS_VARIANTS)))
def test_forest_gibbs_sequential_is_sample_identical(fem, forest_oracle_bin, tmp_path, variant):
    """same uniforms u(seed, sweep, forest, draw) => the same derivation for every forest in every sweep, the same
    per-sweep probabilities and the same final (time-averaged) weights as the CPU restatement (forest.hpp:726-816,
    forest-em.hpp:694-797, gibbs.hpp:803-877)"""
    rng = np.random.default_rng(20262020 + variant)
    d = str(tmp_path)
    open(f"{d}/f", "w").write("\n".join(random_forest(rng, n_rules=40, depth=5, share=0.25) for _ in range(70)) + "\n")
    # (variant 0 leaves some rules out of every group: fixed zero probabilities, cache-model prob = 2^-inf on both sides)
    open(f"{d}/n", "w").write(random_normgroups(rng, 40, leave_out=0.1 if variant == 0 else 0.0))
    args = ["-U", "-f", f"{d}/f", "-n", f"{d}/n", "--seed=5", *GIBBS_VARIANTS[variant]]
    rc, _, oerr = run(forest_oracle_bin, [*args, f"--outsample-file={d}/o.s", "-o", f"{d}/o.w", f"--history={d}/o.h"])
    assert rc == 0, oerr
    rc, _, err = run(fem, [*args, f"--outsample-file={d}/p.s", "-o", f"{d}/p.w", f"--history={d}/p.h"])
    assert rc == 0, err
    assert open(f"{d}/p.s").read() == open(f"{d}/o.s").read()
    ho = [float(ln.split()[1]) for ln in open(f"{d}/o.h") if ln.strip()]
    hp = [float(ln.split()[1]) for ln in open(f"{d}/p.h") if ln.strip()]
    assert len(ho) == len(hp) and len(ho) >= 5
    for a, b in zip(hp, ho):
        assert a == b or abs(a - b) <= 1e-9 * max(1.0, abs(b)), (a, b)
    for x, y in zip(read_weights(f"{d}/p.w"), read_weights(f"{d}/o.w")):
        assert _close_ln(x, y, 1e-8, floor=-600.0), (x, y)
    assert [ln for ln in err.splitlines() if ln.startswith("Gibbs i=")] == [ln for ln in oerr.splitlines() if ln.startswith("Gibbs i=")]


def test_forest_gibbs_converges_to_the_posterior(fem, tmp_path):
    """one forest with two derivations, rule 1 or rule 2, in one group: the sampler is a Polya urn with prior alpha*p0*N
    = 0.5 each and 200 copies of the forest; the time-averaged weight of rule 1 must sit near 1/2, and every sampled
    derivation is one of the two (sanity of the cache / count bookkeeping, no oracle involved)"""
    d = str(tmp_path)
    open(f"{d}/f", "w").write("(OR 1 2)\n" * 200)
    open(f"{d}/n", "w").write("((1 2))\n")
    rc, _, err = run(fem, ["-U", "-f", f"{d}/f", "-n", f"{d}/n", "--crp=400", "--burnin=100", "--const-alpha=0.5", "--seed=9",
                           f"--outsample-file={d}/s", "-o", f"{d}/w"], timeout=600)
    assert rc == 0, err
    w = read_weights(f"{d}/w")
    p1 = math.exp(w[0])
    assert abs(p1 + math.exp(w[1]) - 1) < 1e-9
    assert 0.2 < p1 < 0.8, p1  # (the urn's limit is Beta(0.5, 0.5)-ish per run; the time average pulls it inside)
    assert set(open(f"{d}/s").read().split()) <= {"1", "2"}


def test_forest_gibbs_batched_level(fem, forest_oracle_bin, tmp_path):
    """--crp-batched (every forest against the previous sweep's counts): not the oracle's derivations; its cache-model
    probability settles at the sequential sampler's level"""
    rng = np.random.default_rng(20262121)
    d = str(tmp_path)
    open(f"{d}/f", "w").write("\n".join(random_forest(rng, n_rules=30, depth=4) for _ in range(400)) + "\n")
    open(f"{d}/n", "w").write(random_normgroups(rng, 30, leave_out=0.0))
    args = ["-U", "-f", f"{d}/f", "-n", f"{d}/n", "--seed=5", "--crp=60", "--burnin=30"]
    out = {}
    for name, extra in (("seq", []), ("bat", ["--crp-batched"])):
        rc, _, err = run(fem, [*args, *extra, f"--history={d}/h.{name}"], timeout=600)
        assert rc == 0, err
        h = [float(ln.split()[1]) for ln in open(f"{d}/h.{name}") if ln.strip()]
        out[name] = sum(h[-20:]) / 20
    assert abs(out["seq"] - out["bat"]) <= 0.03 * abs(out["seq"]), out
