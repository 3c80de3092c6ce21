"""Parity of the sparse-emission dense-state kernel (k_fb_sparse: one sequence per lane, <= 8 states per symbol,
final weights) against the CPU oracle (real lattices, carmel/src/derivations.h:400-449,479-704) and against the
product's own lattice path (--no-dense).  HMM / tagging cascades: *e*:tag bigram FSA with an *e*:*e* arc into the
final state, composed with a 1-state tag:word lexicon (carmel-tutorial/tagging.fsa, tagging.fst).

Tolerances (north_star): 1e-6 relative in fp64, 1e-4 in fp32."""
import os
import re

import numpy as np
import pytest

from helpers import compare_wfst_text, read_history, run

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli(native_lib):
    from carmel_b200 import CLI_PATH
    return CLI_PATH


def _close(h_got, h_want, rel):
    assert [h[0] for h in h_got] == [h[0] for h in h_want], (len(h_got), len(h_want))
    for a, b in zip(h_got, h_want):
        assert abs(a[1] - b[1]) <= rel * max(1.0, abs(b[1])), (a, b)
        assert abs(a[2] - b[2]) <= rel * max(1.0, abs(b[2])), (a, b)


CASES = [
    dict(hmm=dict(n_sent=150, n_tags=6, vocab=40, tags_per_word=3, len_range=(1, 30))),            # ragged tiles, K=4
    dict(hmm=dict(n_sent=33, n_tags=8, vocab=20, tags_per_word=4, len_range=(10, 41))),            # one full tile + 1
    dict(hmm=dict(n_sent=40, n_tags=5, vocab=12, tags_per_word=2, len_range=(300, 420))),          # long: exponents
    dict(hmm=dict(n_sent=64, n_tags=10, vocab=200, tags_per_word=1, len_range=(5, 12))),           # 1 state / position
    dict(hmm=dict(n_sent=70, n_tags=40, vocab=60, tags_per_word=6, len_range=(3, 20))),            # K=8, 42 states
    dict(hmm=dict(n_sent=90, n_tags=7, vocab=30, tags_per_word=3, len_range=(2, 25)), lock_fsa=True),   # no xi
    dict(hmm=dict(n_sent=90, n_tags=7, vocab=30, tags_per_word=3, len_range=(2, 25)), lock_lex=True),   # no gamma slots
]


def _make(d, case):
    from carmel_b200 import synth
    w = synth.write_hmm(d, seed=20261300, **case["hmm"])
    data, fsa, fst = w["files"]
    rng = np.random.default_rng(11)
    for path, lock in ((fsa, case.get("lock_fsa")), (fst, case.get("lock_lex"))):
        # random (unnormalised) initial weights so the trajectory is not symmetric; optionally locked
        txt = open(path).read()
        txt = re.sub(r" 1\)\)", lambda m: f" {rng.uniform(0.1, 1.0):.6g}{'!' if lock else ''}))", txt)
        open(path, "w").write(txt)
    return [data, fsa, fst]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("mode,rel", [(["--scaled"], 1e-6), (["--float", "--scaled"], 1e-4)])
def test_sparse_hmm_matches_oracle(cli, oracle_bin, tmp_path, case, mode, rel):
    d = str(tmp_path)
    files = {sub: _make(os.path.join(d, sub), CASES[case]) for sub in ("o", "p", "l")}
    args = ["--train-cascade", "-HJ", "-M", "5"]
    rc, _, oerr = run(oracle_bin, [*args, f"--history={d}/h.o", *files["o"]], timeout=600)
    assert rc == 0, oerr
    rc, _, err = run(cli, [*args, *mode, "--dense", f"--history={d}/h.p", *files["p"]])
    assert rc == 0, err
    assert "one sequence per lane" in err, err
    rc, _, lerr = run(cli, [*args, *mode, "--no-dense", f"--history={d}/h.l", *files["l"]])
    assert rc == 0, lerr
    _close(read_history(f"{d}/h.p"), read_history(f"{d}/h.o"), rel)
    _close(read_history(f"{d}/h.p"), read_history(f"{d}/h.l"), rel)
    for n in ("tags.fsa.trained", "lexicon.fst.trained"):
        compare_wfst_text(open(os.path.join(d, "p", n)).read(), open(os.path.join(d, "o", n)).read(), rel * 20,
                          ln_floor=-690.0 if rel <= 1e-6 else -60.0)


def test_sparse_estep_equals_lattice_estep(native_lib, tmp_path):
    """one E-step through the C ABI on both paths: per-example ln P, the weights after one M-step, and the lattice
    sizes the dense-state path reports without building lattices"""
    import carmel_b200 as cb
    files = _make(str(tmp_path), dict(hmm=dict(n_sent=200, n_tags=7, vocab=60, tags_per_word=3, len_range=(1, 50))))
    out = {}
    for name, extra in (("sparse", ["--dense"]), ("lattice", ["--no-dense"])):
        job = cb.Job(["--train-cascade", "--scaled", "-q", *extra, *files])
        ctx = job.prepare()
        st = job.stats()
        ds = ctx.dense_stats()
        r = ctx.estimate()
        lp = ctx.example_logprob(st["examples"])
        ctx.maximize(1.0)
        out[name] = dict(st=st, ds=ds, sum=(r.sum_ln_p, r.sum_w_ln_p, r.n_zero), lp=lp, w=ctx.get_params())
        job.close()
    a, b = out["sparse"], out["lattice"]
    assert a["ds"]["kernel"] == "sparse" and a["ds"]["k"] == 4 and a["ds"]["sequences"] == 200
    assert b["ds"]["kernel"] is None
    for k in ("examples", "trellis_states", "trellis_arcs"):
        assert a["st"][k] == b["st"][k], (k, a["st"][k], b["st"][k])
    np.testing.assert_allclose(a["lp"], b["lp"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(a["sum"][:2], b["sum"][:2], rtol=1e-10)
    fin = np.isfinite(b["w"])
    assert np.array_equal(fin, np.isfinite(a["w"]))
    np.testing.assert_allclose(a["w"][fin], b["w"][fin], rtol=0, atol=1e-8)


def test_sparse_kernel_without_final_weights(cli, oracle_bin, tmp_path, monkeypatch):
    """a homophonic cipher (every cipher symbol has one or two source letters, the final state is an ordinary
    state): both dense-state kernels apply; CML_SPARSE_MIN_SEQ=1 selects the sparse one for a small corpus"""
    import shutil
    rng = np.random.default_rng(5)
    LET = ["_"] + [chr(65 + i) for i in range(11)]
    CIP = ["_"] + [chr(97 + i) for i in range(18)]
    n = len(LET)
    lm = rng.dirichlet(np.full(n, 0.5), size=n)
    src = {0: [0]}
    for c in range(1, len(CIP)):
        src[c] = sorted(set(int(v) for v in rng.integers(1, n, size=int(rng.integers(1, 3)))))
    q = lambda s_: '"' + s_ + '"'
    d = os.path.join(str(tmp_path), "o")
    os.makedirs(d)
    with open(os.path.join(d, "lm.wfsa"), "w") as f:
        f.write("_\n")
        for a in range(n):
            for b in range(n):
                f.write(f"({LET[a]} ({LET[b]} *e* {q(LET[b])} {lm[a, b]:.12g}))\n")
    with open(os.path.join(d, "channel.fst"), "w") as f:
        f.write("0\n")
        for c in range(len(CIP)):
            for a in src[c]:
                f.write(f"(0 (0 {q(LET[a])} {q(CIP[c])} {rng.uniform(0.2, 1):.5g}))\n")
    with open(os.path.join(d, "c.data"), "w") as f:
        for _ in range(60):
            ln = int(rng.integers(1, 30))
            out = [CIP[int(rng.integers(0, len(CIP)))] for _ in range(ln - 1)] + ["_"]
            f.write("\n" + " ".join(q(c) for c in out) + "\n")
    shutil.copytree(d, os.path.join(str(tmp_path), "p"))
    args = ["--train-cascade", "-HJ", "-M", "6"]
    d = str(tmp_path)
    rc, _, oerr = run(oracle_bin, [*args, f"--history={d}/h.o", "c.data", "lm.wfsa", "channel.fst"], cwd=os.path.join(d, "o"))
    assert rc == 0, oerr
    monkeypatch.setenv("CML_SPARSE_MIN_SEQ", "1")
    rc, _, err = run(cli, [*args, "--scaled", "--dense", f"--history={d}/h.p", "c.data", "lm.wfsa", "channel.fst"],
                     cwd=os.path.join(d, "p"))
    assert rc == 0, err
    assert "one sequence per lane" in err, err
    _close(read_history(f"{d}/h.p"), read_history(f"{d}/h.o"), 1e-6)
    for nme in ("lm.wfsa.trained", "channel.fst.trained"):
        compare_wfst_text(open(os.path.join(d, "p", nme)).read(), open(os.path.join(d, "o", nme)).read(), 2e-5)
