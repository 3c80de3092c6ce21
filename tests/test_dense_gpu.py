"""Parity of the dense-state E-step (cml_add_sequences / k_fb_dense: never-materialised position-synchronous
lattices, w(i->j,o) = T[i][j]*E[j][o]) against the CPU oracle (which builds and walks the real lattices,
carmel/src/derivations.h:400-449,479-704) and against the product's own lattice path (--no-dense).

Tolerances (north_star): likelihood trajectory and learned weights within 1e-6 relative in fp64, 1e-4 in fp32."""
import os

import numpy as np
import pytest

from helpers import compare_wfst_text, read_history, run

pytestmark = pytest.mark.gpu

LET = ["_", "A", "B", "C", "D", "E"]
CIP = ["_", "a", "b", "c", "d", "e", "f"]


@pytest.fixture(scope="module")
def cli(native_lib):
    from carmel_b200 import CLI_PATH
    return CLI_PATH


def q(s):
    return '"' + s + '"'


def write_small_cipher(d, rng, n_lines=30, lock_lm=True, lens=(1, 25), weighted=False, bad_lines=0, sparse_lm=False):
    """letter-bigram LM (*e*:letter FSA over LET, final state '_') o 1-state channel LET x CIP; ciphertext lines end
    in '_' (the final state) except `bad_lines` lines, which have no derivation."""
    n = len(LET)
    lm = rng.dirichlet(np.full(n, 0.5), size=n)
    if sparse_lm:  # some bigrams impossible: the pruned lattice is smaller than the dense one
        lm = lm * (rng.random((n, n)) < 0.6)
        lm[:, 0] = np.maximum(lm[:, 0], 0.05)
        lm[0, :] = np.maximum(lm[0, :], 0.05)
    lm /= lm.sum(1, keepdims=True)
    wfsa, fst, data = (os.path.join(d, f) for f in ("lm.wfsa", "channel.fst", "cipher.data"))
    with open(wfsa, "w") as f:
        f.write("_\n")
        for a in range(n):
            for b in range(n):
                if lm[a, b] > 0:
                    f.write(f"({LET[a]} ({LET[b]} *e* {q(LET[b])} {lm[a, b]:.12g}{'!' if lock_lm else ''}))\n")
    with open(fst, "w") as f:
        f.write("0\n")
        for a in range(n):
            for b in range(len(CIP)):
                f.write(f"(0 (0 {q(LET[a])} {q(CIP[b])} {rng.uniform(0.2, 1.0):.6g}))\n")
    with open(data, "w") as f:
        for k in range(n_lines):
            ln = int(rng.integers(lens[0], lens[1] + 1))
            out = [CIP[int(rng.integers(0, len(CIP)))] for _ in range(ln - 1)] + ["_"]
            if k < bad_lines:
                out[-1] = "zz"  # unknown symbol -> no derivation
            if weighted:
                f.write(f"{rng.uniform(0.5, 3.0):.4g}\n")
            f.write("\n" + " ".join(q(c) for c in out) + "\n")
    return data, wfsa, fst


def _close(h_got, h_want, rel):
    assert [h[0] for h in h_got] == [h[0] for h in h_want], (len(h_got), len(h_want))
    for a, b in zip(h_got, h_want):
        assert abs(a[1] - b[1]) <= rel * max(1.0, abs(b[1])), (a, b)
        assert abs(a[2] - b[2]) <= rel * max(1.0, abs(b[2])), (a, b)


CASES = [
    dict(),                                   # locked LM: gamma counts only
    dict(lock_lm=False),                      # trainable transitions: xi counts
    dict(lock_lm=False, weighted=True),       # example weights
    dict(bad_lines=3),                        # examples without a derivation are dropped like the lattice builder does
    dict(sparse_lm=True, lock_lm=False),      # impossible bigrams: dead / unreachable lattice states
    dict(lens=(1, 3), n_lines=50),            # very short lines (a 1-letter line has one lattice arc)
    dict(lens=(90, 130), n_lines=8),          # lines longer than the 32-symbol prefetch chunks
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("mode,rel", [(["--scaled"], 1e-6), (["--float", "--scaled"], 1e-4)])
def test_dense_em_matches_oracle(cli, oracle_bin, tmp_path, case, mode, rel):
    rng = np.random.default_rng(20261001 + case)
    for sub in ("o", "p", "l"):
        os.makedirs(os.path.join(str(tmp_path), sub))
    files = {}
    for sub in ("o", "p", "l"):
        files[sub] = write_small_cipher(os.path.join(str(tmp_path), sub), np.random.default_rng(20261001 + case), **CASES[case])
    args = ["--train-cascade", "-HJ", "-M", "8"]
    d = str(tmp_path)
    rc, _, oerr = run(oracle_bin, [*args, f"--history={d}/h.o", *files["o"]])
    assert rc == 0, oerr
    rc, _, err = run(cli, [*args, *mode, f"--history={d}/h.p", *files["p"]])
    assert rc == 0, err
    assert "dense-state path: 6 states x 7 symbols" in err, err
    rc, _, lerr = run(cli, [*args, *mode, "--no-dense", f"--history={d}/h.l", *files["l"]])
    assert rc == 0, lerr
    assert "dense-state path" not in lerr
    assert err.count("No derivations in transducer") == oerr.count("No derivations in transducer") == CASES[case].get("bad_lines", 0)
    _close(read_history(f"{d}/h.p"), read_history(f"{d}/h.o"), rel)
    _close(read_history(f"{d}/h.p"), read_history(f"{d}/h.l"), rel)
    for n in ("lm.wfsa.trained", "channel.fst.trained"):
        compare_wfst_text(open(os.path.join(d, "p", n)).read(), open(os.path.join(d, "o", n)).read(), rel * 20,
                          ln_floor=-690.0 if rel <= 1e-6 else -60.0)


def test_dense_estep_equals_lattice_estep(native_lib, tmp_path):
    """one E-step through the C ABI on both paths: per-example ln P, corpus sums, the weights after one M-step,
    and the lattice sizes the dense path reports (it never builds the lattices) equal the lattice builder's"""
    import carmel_b200 as cb
    rng = np.random.default_rng(7)
    files = write_small_cipher(str(tmp_path), rng, n_lines=40, lock_lm=False, sparse_lm=True, weighted=True)
    out = {}
    for name, extra in (("dense", []), ("lattice", ["--no-dense"])):
        job = cb.Job(["--train-cascade", "--scaled", "-q", *extra, *files])
        ctx = job.prepare()
        st = job.stats()
        r = ctx.estimate()
        lp = ctx.example_logprob(st["examples"])
        ctx.maximize(1.0)
        out[name] = dict(st=st, sum=(r.sum_ln_p, r.sum_w_ln_p, r.n_zero), lp=lp, w=ctx.get_params(),
                         dense=ctx.dense_stats())
        job.close()
    a, b = out["dense"], out["lattice"]
    assert a["dense"]["sequences"] == 40 and a["dense"]["t_slots"] > 0 and a["dense"]["e_slots"] == 42
    assert b["dense"]["sequences"] == 0
    for k in ("examples", "trellis_states", "trellis_arcs"):
        assert a["st"][k] == b["st"][k], (k, a["st"][k], b["st"][k])
    assert a["sum"][2] == b["sum"][2] == 0
    np.testing.assert_allclose(a["lp"], b["lp"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(a["sum"][:2], b["sum"][:2], rtol=1e-10)
    fin = np.isfinite(b["w"])
    assert np.array_equal(fin, np.isfinite(a["w"]))
    np.testing.assert_allclose(a["w"][fin], b["w"][fin], rtol=0, atol=1e-8)


def test_dense_zero_probability_sequence(native_lib):
    """a sequence whose only derivation uses a zero-weight arc has P = 0: reported in n_zero, contributes no counts
    (train.cc:326-332 / the lattice kernels do the same)"""
    import carmel_b200 as cb
    # 2 states, 2 symbols; T cells are parameters 0..3 (i*2+j), E cells parameters 4..7 (j*2+o); arc (i,j,o) = T*E
    arcs = [(i, j, o) for i in range(2) for j in range(2) for o in range(2)]
    chain_off = np.arange(0, 2 * len(arcs) + 1, 2, dtype=np.uint32)
    chain = np.asarray([[i * 2 + j, 4 + j * 2 + o] for (i, j, o) in arcs], np.uint32).ravel()
    ctx = cb.Context(0, 64, cb.SPACE_SCALED)
    grp = np.asarray([0, 0, 1, 1, 2, 2, 3, 3], np.uint32)
    ctx.set_model(len(arcs), 8, grp, np.full(8, cb.NO_GROUP, np.uint32), 4, chain_off=chain_off, chain_param=chain)
    with np.errstate(divide="ignore"):
        ctx.set_params(np.log(np.asarray([0.5, 0.5, 0.5, 0.5, 1.0, 0.0, 0.3, 0.7])))  # E[0][1] = 0
    src, dst, sym = (np.asarray([a[k] for a in arcs], np.uint32) for k in range(3))
    seqs = [[0, 0, 0], [1, 1, 0], [], [1, 0]]  # final state 0: [1,1,0] must end in state 0 emitting 0: fine
    ctx.add_sequences(2, 2, 0, 0, src, dst, sym, seqs)
    r = ctx.estimate()
    lp = ctx.example_logprob(4)
    # brute force
    T = np.full((2, 2), 0.5)
    E = np.asarray([[1.0, 0.0], [0.3, 0.7]])
    want = []
    for s in seqs:
        a = np.asarray([1.0, 0.0])
        for o in s:
            a = (a @ T) * E[:, o]
        want.append(a[0])
    for g, p in zip(lp, want):
        if p == 0:
            assert g == -np.inf
        else:
            assert abs(g - np.log(p)) < 1e-12
    assert r.n_zero == sum(1 for p in want if p == 0)
    ctx.close()


def test_not_dense_falls_back(cli, oracle_bin, tmp_path):
    """every arc consumes one output symbol but each arc is its own parameter (no T x E factorisation):
    the library answers CML_ERR_NOT_DENSE and the job trains on lattices, same results as the oracle"""
    rng = np.random.default_rng(3)
    f, c = os.path.join(str(tmp_path), "h.fst"), os.path.join(str(tmp_path), "h.data")
    with open(f, "w") as fh:
        fh.write("s0\n")
        for i in range(3):
            for j in range(3):
                for o in "xyz":
                    fh.write(f"(s{i} (s{j} *e* {q(o)} {rng.uniform(0.1, 1):.5g}))\n")
    with open(c, "w") as fh:
        for _ in range(12):
            fh.write("\n" + " ".join(q("xyz"[int(rng.integers(0, 3))]) for _ in range(int(rng.integers(1, 9)))) + "\n")
    rc, oout, oerr = run(oracle_bin, ["-t", "-M", "6", f"--history={tmp_path}/h.o", c, f])
    assert rc == 0, oerr
    rc, out, err = run(cli, ["-t", "-M", "6", "--scaled", f"--history={tmp_path}/h.p", c, f])
    assert rc == 0, err
    assert "dense-state view not applicable" in err
    _close(read_history(f"{tmp_path}/h.p"), read_history(f"{tmp_path}/h.o"), 1e-6)
    compare_wfst_text(out, oout, 1e-5)
    rc, _, err = run(cli, ["-t", "-M", "6", "--scaled", "--dense", c, f])
    assert rc != 0 and "no dense-state view" in err


@pytest.mark.parametrize("mode,rel", [(["--scaled"], 1e-6), (["--float", "--scaled"], 1e-4)])
def test_synthetic_27_state_cipher_all_paths(cli, oracle_bin, tmp_path, mode, rel):
    """the bench's cipher model at a small size: dense-state kernel, and with --no-dense the 32x8 one-CTA-per-lattice
    class of the level-sliced ELL kernel (27-state x 27-arc levels, aggregated count columns), both against the oracle"""
    from carmel_b200 import synth
    d = str(tmp_path)
    files = {sub: synth.write_cipher(os.path.join(d, sub), n_lines=24, line_len=14, seed=77)["files"] for sub in ("o", "p", "l")}
    args = ["--train-cascade", "-HJ", "-M", "5"]
    rc, _, oerr = run(oracle_bin, [*args, f"--history={d}/h.o", *files["o"]], timeout=600)
    assert rc == 0, oerr
    rc, _, err = run(cli, [*args, *mode, "--dense", f"--history={d}/h.p", *files["p"]])
    assert rc == 0, err
    rc, _, lerr = run(cli, [*args, *mode, "--no-dense", f"--history={d}/h.l", *files["l"]])
    assert rc == 0, lerr
    _close(read_history(f"{d}/h.p"), read_history(f"{d}/h.o"), rel)
    _close(read_history(f"{d}/h.l"), read_history(f"{d}/h.o"), rel)
    for sub in ("p", "l"):
        compare_wfst_text(open(os.path.join(d, sub, "channel.fst.trained")).read(),
                          open(os.path.join(d, "o", "channel.fst.trained")).read(), rel * 20,
                          ln_floor=-690.0 if rel <= 1e-6 else -60.0)


def test_dense_long_lines(cli, oracle_bin, tmp_path):
    """SURVEY 8(d) C2 also names the one-long-line variant of the cipher corpus: two 3,000-letter lines (alpha / beta are
    renormalised thousands of times; the likelihood is ~2^-14000), dense-state kernel against the oracle's lattices"""
    from carmel_b200 import synth
    d = str(tmp_path)
    files = {sub: synth.write_cipher(os.path.join(d, sub), n_lines=2, line_len=3000, seed=5)["files"] for sub in ("o", "p")}
    args = ["--train-cascade", "-HJ", "-M", "3"]
    rc, _, oerr = run(oracle_bin, [*args, f"--history={d}/h.o", *files["o"]], timeout=900)
    assert rc == 0, oerr
    for mode, rel in ((["--scaled"], 1e-6), (["--float", "--scaled"], 1e-4)):
        rc, _, err = run(cli, [*args, *mode, "--dense", f"--history={d}/h.p", *files["p"]])
        assert rc == 0, err
        _close(read_history(f"{d}/h.p"), read_history(f"{d}/h.o"), rel)


@pytest.mark.parametrize("lock_lm", [True, False])
def test_tensor_core_sweeps_match_oracle(cli, oracle_bin, tmp_path, monkeypatch, lock_lm):
    """k_dense_tc (3xTF32 mma.sync, 16 sequences per warp; normally for >= 16,384 sequences, forced here with
    CML_DENSE_TC=1): fp32 trajectory and weights against the oracle on ragged lines (rows of a 16-sequence group
    start and end at different positions) and a partly filled last group; lock_lm=False adds the transition counts
    xi as a third tensor-core product per position ([32 x 16] . [16 x 32], K = the warp's 16 sequences)"""
    monkeypatch.setenv("CML_DENSE_TC", "1")
    d = str(tmp_path)
    rng = np.random.default_rng(99)
    files = {}
    for sub in ("o", "p"):
        os.makedirs(os.path.join(d, sub))
        files[sub] = write_small_cipher(os.path.join(d, sub), np.random.default_rng(99), n_lines=53, lens=(1, 70), weighted=True,
                                        lock_lm=lock_lm)
    del rng
    args = ["--train-cascade", "-HJ", "-M", "6"]
    rc, _, oerr = run(oracle_bin, [*args, f"--history={d}/h.o", *files["o"]])
    assert rc == 0, oerr
    rc, _, err = run(cli, [*args, "--float", "--scaled", "--dense", f"--history={d}/h.p", *files["p"]])
    assert rc == 0, err
    assert "3xTF32 tensor-core sweeps" in err, err
    _close(read_history(f"{d}/h.p"), read_history(f"{d}/h.o"), 1e-4)
    for name in ("channel.fst.trained", "lm.wfsa.trained"):
        compare_wfst_text(open(os.path.join(d, "p", name)).read(), open(os.path.join(d, "o", name)).read(), 2e-3,
                          ln_floor=-60.0)
