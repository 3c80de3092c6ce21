"""`--crp` Gibbs sampling parity on the GPU.

Sequential mode: product and oracle draw the same counter-based uniforms (seed, sweep, block, draw),
so every sampled derivation must be IDENTICAL arc for arc, the per-sweep cache-model probability must
agree to 1e-9 relative and the final (time-averaged) weights to 1e-6.  Batched mode is a different
(parallel, stale-count) sampler: it is checked for self-consistency only (valid paths, finite
probability, normalised weights)."""
import math
import os
import re

import numpy as np
import pytest

from helpers import _ln_weight, compare_wfst_text, random_wfst, run, sample_pairs, stage

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli(native_lib):
    from carmel_b200 import CLI_PATH
    return CLI_PATH


def _hist(path):
    return [(int(r.split()[0]), float(r.split()[1])) for r in open(path) if r.strip()]


def _samples(path):
    return [ln.split() for ln in open(path)]


def _both(cli, oracle_bin, tmp_path, args, files, trained):
    rc, out, err = run(cli, [*args, f"--history={tmp_path}/h.p", f"--dump-samples={tmp_path}/s.p", *files])
    assert rc == 0, err
    got_trained = [open(f + ".trained").read() for f in trained]
    rc, oout, oerr = run(oracle_bin, [*args, f"--history={tmp_path}/h.o", f"--dump-samples={tmp_path}/s.o", *files])
    assert rc == 0, oerr
    want_trained = [open(f + ".trained").read() for f in trained]
    hp, ho = _hist(f"{tmp_path}/h.p"), _hist(f"{tmp_path}/h.o")
    assert len(hp) == len(ho)
    for (i, a), (j, b) in zip(hp, ho):
        assert i == j and abs(a - b) <= 1e-9 * max(1.0, abs(b)), (i, a, b)
    sp, so = _samples(f"{tmp_path}/s.p"), _samples(f"{tmp_path}/s.o")
    assert sp == so
    for g, w in zip(got_trained, want_trained):
        compare_wfst_text(g, w, 1e-6)
    return err


@pytest.mark.parametrize("extra", [[], ["--burnin=3"], ["--final-counts"], ["--crp-exclude-prior"], ["--uniform-p0"],
                                   ["--dirichlet-p0"], ["--high-temp=3", "--low-temp=1"]])
def test_cipher_crp_sequential(cli, oracle_bin, tmp_path, extra):
    data, wfsa, fst = stage(tmp_path, "cipher.data", "cipher.wfsa", "cipher.fst")
    err = _both(cli, oracle_bin, tmp_path, ["--crp", "-M", "8", "--priors=0,1e-2", "--seed=7", "-HJ", *extra],
                [data, wfsa, fst], [wfsa, fst])
    assert "Gibbs i=8" in err


def test_tagging_crp_sequential(cli, oracle_bin, tmp_path):
    data, fsa, fst = stage(tmp_path, "tagging.data", "tagging.fsa", "tagging.fst")
    _both(cli, oracle_bin, tmp_path, ["--crp=5", "--burnin=2", "--priors=0.1,0.01", "--seed=3", "-HJ"], [data, fsa, fst],
          [fsa, fst])


def test_epron_crp_single_transducer(cli, oracle_bin, tmp_path):
    fst, data = stage(tmp_path, "epron-jpron.fst", "epron-jpron.data")
    _both(cli, oracle_bin, tmp_path, ["--crp", "-M", "6", "--priors=0.05", "--seed=11"], [data, fst], [fst])


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_crp_sequential(cli, oracle_bin, tmp_path, seed):
    rng = np.random.default_rng(1000 + seed)
    ns = int(rng.integers(3, 7))
    text, ins, outs, arcs = random_wfst(rng, n_states=ns, eps_rate=0.15, lock_rate=0.1, tie_rate=0.0)
    fst = os.path.join(str(tmp_path), "r.fst")
    open(fst, "w").write(text)
    data = os.path.join(str(tmp_path), "r.data")
    open(data, "w").write(sample_pairs(rng, arcs, ns, n_pairs=12, noise=0.0))
    _both(cli, oracle_bin, tmp_path, ["--crp", "-M", "5", "--priors=0.02", f"--seed={seed}"], [data, fst], [fst])


def test_batched_final_perplexity_agrees_with_sequential(cli, tmp_path):
    """north_star: batched Gibbs is reported as final-perplexity agreement with the exact sequential sampler"""
    # Batched sweeps sample every block against the previous sweep's counts minus the block's own previous sample
    # (a synchronous, stale-count version of the exact sampler).
    # (1) tutorial cipher, 10 blocks: per-symbol cache-model perplexity of the two samplers after 300 sweeps
    data, wfsa, fst = stage(tmp_path, "cipher.data", "cipher.wfsa", "cipher.fst")
    tail = {}
    for name, extra in (("seq", []), ("bat", ["--crp-batched"])):
        rc, out, err = run(cli, ["--crp", "-M", "300", "--burnin=100", *extra, "--priors=0,1e-2", "--seed=9", "-q",
                                 f"--history={tmp_path}/h.{name}", data, wfsa, fst], timeout=600)
        assert rc == 0, err
        h = _hist(f"{tmp_path}/h.{name}")
        tail[name] = np.mean([v for _, v in h[-100:]])  # ln cache-model probability, averaged over the last 100 sweeps
    ppx = {k: -v / 505 / math.log(2) for k, v in tail.items()}  # 505 output symbols (commands.trace per-point-ppx N)
    assert abs(ppx["seq"] - ppx["bat"]) <= 0.08 * ppx["seq"], ppx
    # (2) synthetic cipher, 300 blocks x 30 letters: the batched sampler must end near the EM optimum of the same model
    # (the exact sampler mixes slowly on ciphers -- the reference's own tutorial runs it for 6000 sweeps -- and is still
    # far above it after a few hundred sweeps, so it is only required not to beat the batched result by much)
    from carmel_b200 import synth
    w = synth.write_cipher(str(tmp_path / "syn"), n_lines=300, line_len=30, seed=7)
    data, wfsa, fst = w["files"]
    rc, out, err = run(cli, ["--train-cascade", "-M", "60", data, wfsa, fst], timeout=600)
    assert rc == 0, err
    em_bits = float(re.findall(r"per-output-symbol-perplexity\(N=9000\)=2\^([0-9.]+)", err)[-1])
    rc, out, err = run(cli, ["--crp", "-M", "150", "--burnin=50", "--crp-batched", "--priors=0,1e-2", "--seed=9", "-q",
                             f"--history={tmp_path}/h.syn", data, wfsa, fst], timeout=600)
    assert rc == 0, err
    bat_bits = -np.mean([v for _, v in _hist(f"{tmp_path}/h.syn")[-50:]]) / 9000 / math.log(2)
    assert abs(bat_bits - em_bits) <= 0.05 * em_bits, (bat_bits, em_bits)


def test_cipher_crp_batched(cli, tmp_path):
    data, wfsa, fst = stage(tmp_path, "cipher.data", "cipher.wfsa", "cipher.fst")
    rc, out, err = run(cli, ["--crp", "-M", "10", "--crp-batched", "--priors=0,1e-2", "--seed=5", "-HJ",
                             f"--history={tmp_path}/h.p", f"--dump-samples={tmp_path}/s.p", data, wfsa, fst])
    assert rc == 0, err
    h = _hist(f"{tmp_path}/h.p")
    assert len(h) == 11 and all(np.isfinite(v) and v < 0 for _, v in h)
    # the sampler must move towards better samples than the initial (prior) draw
    assert max(v for _, v in h[5:]) > h[0][1]
    # every block sampled a path of the letter count of its line
    lens = [len(s) for s in _samples(f"{tmp_path}/s.p")]
    assert all(n > 0 for n in lens)
    # trained channel rows are normalised
    rows = {}
    for m in re.finditer(r'\(0 \(0 ("[^"]*") ("[^"]*") ([^)\s]+)\)\)', open(fst + ".trained").read()):
        rows[m.group(1)] = rows.get(m.group(1), 0.0) + math.exp(_ln_weight(m.group(3)))
    assert len(rows) == 27
    for k, v in rows.items():
        assert abs(v - 1) < 1e-6, (k, v)


def _sample_histogram(native_lib, files, extra, n_draws):
    """n_draws independent samples of every block from FIXED weights (init_from_params: the sampler ignores the counts),
    as a histogram over the sampled arc-id tuples of block 0"""
    import ctypes as C
    import carmel_b200 as cb
    job = cb.Job(["--crp", "--crp-batched", "--priors=0,1e-2", "-q", *extra, *files])
    ctx = job.prepare()
    st = job.stats()
    cap = ctx.gibbs_sample_capacity()
    lens = np.zeros(st["examples"], np.uint32)
    arcs = np.zeros(cap, np.uint32)
    hist = {}
    for k in range(n_draws):
        ctx.gibbs_sweep(1, k, seed=123, power=1.0, init_from_params=True)
        ctx.gibbs_get_samples_ptr(lens.ctypes.data, arcs.ctypes.data, cap)
        key = tuple(int(v) for v in arcs[:lens[0]])
        hist[key] = hist.get(key, 0) + 1
    dense = st["dense"]
    job.close()
    return hist, dense


def test_dense_state_sampler_draws_from_the_lattice_samplers_distribution(native_lib, tmp_path):
    """cml_gibbs_attach_dense (batched sweeps on position-synchronous lattices) samples the same posterior over
    derivations as the lattice sampler: with fixed weights every sweep is an independent draw, so the two histograms
    over derivations of one short block must agree (two-sample chi-square on the frequent derivations)"""
    rng = np.random.default_rng(4)
    LET, CIP = ["_", "A", "B", "C"], ["_", "a", "b", "c", "d"]
    d = str(tmp_path)
    q = lambda s_: '"' + s_ + '"'
    lm = rng.dirichlet(np.full(4, 0.7), size=4)
    with open(f"{d}/lm.wfsa", "w") as f:
        f.write("_\n")
        for a in range(4):
            for b in range(4):
                f.write(f"({LET[a]} ({LET[b]} *e* {q(LET[b])} {lm[a, b]:.12g}!))\n")
    with open(f"{d}/ch.fst", "w") as f:
        f.write("0\n")
        for a in range(4):
            for b in range(5):
                f.write(f"(0 (0 {q(LET[a])} {q(CIP[b])} {rng.uniform(0.1, 1):.6g}))\n")
    with open(f"{d}/c.data", "w") as f:
        f.write("\n" + " ".join(q(c) for c in ["a", "c", "b", "d", "_"]) + "\n")
        f.write("\n" + " ".join(q(c) for c in ["b", "_"]) + "\n")
    files = [f"{d}/c.data", f"{d}/lm.wfsa", f"{d}/ch.fst"]
    n = 6000
    hd, dense_d = _sample_histogram(native_lib, files, [], n)
    hl, dense_l = _sample_histogram(native_lib, files, ["--no-dense"], n)
    assert dense_d == 1 and dense_l == 0
    assert all(len(k) == 5 for k in hd) and all(len(k) == 5 for k in hl)
    keys = [k for k in set(hd) | set(hl) if hd.get(k, 0) + hl.get(k, 0) >= 40]
    assert len(keys) >= 8
    chi2 = sum((hd.get(k, 0) - hl.get(k, 0)) ** 2 / (hd.get(k, 0) + hl.get(k, 0)) for k in keys)
    # chi-square with len(keys)-1 degrees of freedom: mean df, sd sqrt(2 df); 6 sd is a comfortable bound
    assert chi2 <= len(keys) + 6 * math.sqrt(2 * len(keys)), (chi2, len(keys))
    # and the derivations themselves are the same set (no derivation one sampler can draw and the other cannot)
    assert {k for k in hd if hd[k] >= 40} == {k for k in hl if hl[k] >= 40}


# ---- --expectation (row a18): blocks carry the posteriors of all their arcs (incremental EM over the CRP counts) -----
def _expect_both(cli, oracle_bin, tmp_path, args, files, trained):
    rc, out, err = run(cli, [*args, "--expectation", f"--history={tmp_path}/h.p", *files])
    assert rc == 0, err
    got = [open(f + ".trained").read() for f in trained]
    rc, oout, oerr = run(oracle_bin, [*args, "--expectation", f"--history={tmp_path}/h.o", *files])
    assert rc == 0, oerr
    want = [open(f + ".trained").read() for f in trained]
    hp, ho = _hist(f"{tmp_path}/h.p"), _hist(f"{tmp_path}/h.o")
    assert len(hp) == len(ho) >= 4
    for (i, a), (j, b) in zip(hp, ho):
        assert i == j and abs(a - b) <= 1e-9 * max(1.0, abs(b)), (i, a, b)
    assert hp[-1][1] > hp[0][1]  # the likelihood of the corpus improves from sweep to sweep
    for g, w in zip(got, want):
        compare_wfst_text(g, w, 1e-6)
    assert "sum-all-derivations prob=" in err
    return err


@pytest.mark.parametrize("extra", [[], ["--burnin=2"], ["--final-counts"], ["--uniform-p0"]])
def test_cipher_crp_expectation(cli, oracle_bin, tmp_path, extra):
    """gibbs.cc:311-316 + derivations.h:381-398 (collect_counts_gibbs): per-sweep sum-all-derivations probability and the
    final weights against the CPU restatement (no random numbers are involved: the comparison is exact up to rounding)"""
    data, wfsa, fst = stage(tmp_path, "cipher.data", "cipher.wfsa", "cipher.fst")
    _expect_both(cli, oracle_bin, tmp_path, ["--crp", "-M", "6", "--priors=0,1e-2", "-HJ", *extra], [data, wfsa, fst], [wfsa, fst])


def test_epron_crp_expectation(cli, oracle_bin, tmp_path):
    fst, data = stage(tmp_path, "epron-jpron.fst", "epron-jpron.data")
    _expect_both(cli, oracle_bin, tmp_path, ["--crp", "-M", "6", "--priors=0.05"], [data, fst], [fst])


def test_tagging_crp_expectation(cli, oracle_bin, tmp_path):
    data, fsa, fst = stage(tmp_path, "tagging.data", "tagging.fsa", "tagging.fst")
    _expect_both(cli, oracle_bin, tmp_path, ["--crp=4", "--burnin=1", "--priors=0.1,0.01", "-HJ"], [data, fsa, fst], [fsa, fst])


def test_batched_sweeps_two_gpus_agree_with_one(cli, tmp_path):
    """batched sweeps shard the blocks over the GPUs, one all-reduce of the count deltas per sweep; the
    sharded run must end at the perplexity of the one-GPU run (not the same samples: the ranks draw different uniforms)
    and write a normalised channel"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from carmel_b200 import synth
    w = synth.write_cipher(str(tmp_path / "syn"), n_lines=300, line_len=30, seed=7)
    data, wfsa, fst = w["files"]
    bits = {}
    for name, extra in (("one", []), ("two", ["--gpus=2"])):
        rc, out, err = run(cli, ["--crp", "-M", "150", "--burnin=50", "--crp-batched", "--priors=0,1e-2", "--seed=9", "-q", "-HJ", *extra,
                                 f"--history={tmp_path}/h.{name}", data, wfsa, fst], timeout=240)
        assert rc == 0, err
        h = _hist(f"{tmp_path}/h.{name}")
        assert len(h) == 151
        bits[name] = -np.mean([v for _, v in h[-50:]]) / 9000 / math.log(2)
        rows = {}
        for m in re.finditer(r'\(0 \(0 ("[^"]*") ("[^"]*") ([^)\s]+)\)\)', open(fst + ".trained").read()):
            rows[m.group(1)] = rows.get(m.group(1), 0.0) + math.exp(_ln_weight(m.group(3)))
        assert len(rows) == 27 and all(abs(v - 1) < 1e-6 for v in rows.values()), (name, rows)
    assert abs(bits["one"] - bits["two"]) <= 0.03 * bits["one"], bits
    # the exact sampler is sequential over the corpus: refused, not silently replicated
    rc, _, err = run(cli, ["--crp", "-M", "5", "--gpus=2", data, wfsa, fst])
    assert rc != 0 and "sequential over the corpus" in err
