"""Parity of the lane-per-lattice E-step kernel (k_fb_lane: tiles of 32 narrow lattices, streams aligned by state
ordinal) against the CPU oracle and against the group kernels (--no-lane) on the same inputs.  The kernel is
normally used for batches of >= 16384 eligible lattices; the tests force it with --lane-min=1 (and --no-dense: the
HMM cascades here also have a dense-state view, which the product would otherwise prefer).

Tolerances (north_star): 1e-6 relative in fp64, 1e-4 in fp32."""
import os

import numpy as np
import pytest

from helpers import compare_wfst_text, random_wfst, read_history, run, sample_pairs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli(native_lib):
    from carmel_b200 import CLI_PATH
    return CLI_PATH


def _close(h_got, h_want, rel):
    assert [h[0] for h in h_got] == [h[0] for h in h_want], (len(h_got), len(h_want))
    for a, b in zip(h_got, h_want):
        if a[1] == b[1] and a[2] == b[2]:
            continue
        assert abs(a[1] - b[1]) <= rel * max(1.0, abs(b[1])), (a, b)
        assert abs(a[2] - b[2]) <= rel * max(1.0, abs(b[2])), (a, b)


HMM_CASES = [
    dict(n_sent=150, n_tags=6, vocab=40, tags_per_word=3, len_range=(1, 30)),     # ragged, 5 partly filled tiles
    dict(n_sent=33, n_tags=8, vocab=20, tags_per_word=4, len_range=(10, 41)),     # one full tile + 1 lattice
    dict(n_sent=40, n_tags=5, vocab=12, tags_per_word=2, len_range=(300, 420)),   # long: per-level rescaling in fp32 and fp64
    dict(n_sent=64, n_tags=10, vocab=200, tags_per_word=1, len_range=(5, 12)),    # unambiguous words: 1 state per level
]


@pytest.mark.parametrize("case", range(len(HMM_CASES)))
@pytest.mark.parametrize("mode,rel", [(["--scaled"], 1e-6), (["--float", "--scaled"], 1e-4)])
def test_lane_hmm_matches_oracle(cli, oracle_bin, tmp_path, case, mode, rel):
    from carmel_b200 import synth
    d = str(tmp_path)
    files = {}
    for sub in ("o", "p", "g"):
        files[sub] = synth.write_hmm(os.path.join(d, sub), seed=20261100 + case, **HMM_CASES[case])["files"]
    args = ["--train-cascade", "-HJ", "-M", "5"]
    rc, _, oerr = run(oracle_bin, [*args, f"--history={d}/h.o", *files["o"]], timeout=600)
    assert rc == 0, oerr
    rc, _, err = run(cli, [*args, *mode, "--no-dense", "--lane-min=1", f"--history={d}/h.p", *files["p"]])
    assert rc == 0, err
    assert "dense-state path" not in err
    rc, _, gerr = run(cli, [*args, *mode, "--no-dense", "--no-lane", f"--history={d}/h.g", *files["g"]])
    assert rc == 0, gerr
    _close(read_history(f"{d}/h.p"), read_history(f"{d}/h.o"), rel)
    _close(read_history(f"{d}/h.p"), read_history(f"{d}/h.g"), rel)
    for n in ("tags.fsa.trained", "lexicon.fst.trained"):
        compare_wfst_text(open(os.path.join(d, "p", n)).read(), open(os.path.join(d, "o", n)).read(), rel * 20,
                          ln_floor=-690.0 if rel <= 1e-6 else -60.0)


def test_lane_layout_is_used_and_equals_group_estep(native_lib, tmp_path):
    """one E-step through the C ABI: per-example ln P and the parameters after one M-step, lane vs group kernels"""
    import carmel_b200 as cb
    from carmel_b200 import synth
    w = synth.write_hmm(str(tmp_path), n_sent=200, n_tags=7, vocab=60, tags_per_word=3, seed=5, len_range=(1, 50))
    out = {}
    for name, extra in (("lane", ["--lane-min=1"]), ("group", ["--no-lane"])):
        job = cb.Job(["--scaled", "--no-dense", "-q", *extra, *w["argv"]])
        ctx = job.prepare()
        st = job.stats()
        ls = ctx.lane_stats()
        r = ctx.estimate()
        lp = ctx.example_logprob(st["examples"])
        ctx.maximize(1.0)
        out[name] = dict(st=st, ls=ls, sum=(r.sum_ln_p, r.sum_w_ln_p, r.n_zero), lp=lp, w=ctx.get_params())
        job.close()
    a, b = out["lane"], out["group"]
    assert a["ls"]["lane_examples"] == 200 and a["ls"]["tiles"] == 7 and a["ls"]["lane_arcs"] == a["st"]["trellis_arcs"]
    assert b["ls"]["lane_examples"] == 0
    np.testing.assert_allclose(a["lp"], b["lp"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(a["sum"][:2], b["sum"][:2], rtol=1e-10)
    fin = np.isfinite(b["w"])
    assert np.array_equal(fin, np.isfinite(a["w"]))
    np.testing.assert_allclose(a["w"][fin], b["w"][fin], rtol=0, atol=1e-8)


@pytest.mark.parametrize("seed", range(8))
def test_lane_random_models_match_oracle(cli, oracle_bin, tmp_path, seed):
    """seeded random transducers: batches mix lane-eligible lattices with ones that need the group / CSR kernels
    (epsilon arcs that skip levels, wide levels); locked and tied arcs, priors"""
    rng = np.random.default_rng(20261200 + seed)
    ns = int(rng.integers(2, 5))
    fst, ins, outs, arcs = random_wfst(rng, n_states=ns, eps_rate=0.15, lock_rate=0.15, tie_rate=0.2)
    corpus = sample_pairs(rng, arcs, ns, n_pairs=40, weighted=bool(seed % 2), ins=ins, outs=outs)
    f, c = os.path.join(str(tmp_path), "r.fst"), os.path.join(str(tmp_path), "r.data")
    open(f, "w").write(fst)
    open(c, "w").write(corpus)
    opts = [[], ["-j"], ["-f", "0.01"], ["--priors=0.1"]][seed % 4]
    args = ["-t", "-M", "8", *opts]
    rc, oout, oerr = run(oracle_bin, [*args, f"--history={tmp_path}/h.o", c, f])
    assert rc == 0, oerr
    if any(abs(h[1]) < 1e-6 for h in read_history(f"{tmp_path}/h.o")):
        pytest.skip("degenerate corpus (probability 1)")
    rc, out, err = run(cli, [*args, "--scaled", "--no-dense", "--lane-min=1", f"--history={tmp_path}/h.p", c, f])
    assert rc == 0, err
    _close(read_history(f"{tmp_path}/h.p"), read_history(f"{tmp_path}/h.o"), 1e-6)
    compare_wfst_text(out, oout, 1e-5)
