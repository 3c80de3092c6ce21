"""Round-2 parity tests through the C ABI / the command lines on the GPU:

  * the fused one-synchronisation EM step (cml_em_step, CUDA graph) against the two-call path;
  * an N = 2 sharded E-step THROUGH THE REAL KERNELS on one GPU (two contexts, reduce buffers summed like the all-reduce
    does) against the unsharded E-step: sum ln P and every count slot;
  * every kernel that is selected automatically by a size threshold, at a size that crosses the threshold WITHOUT any
    forcing option (lane kernel >= 16384 lattices, 3xTF32 dense sweeps >= 16384 fp32 sequences, thread-per-forest tiles
    >= 8192 forests), against the CPU oracle;
  * `-M 0 --train-cascade` (fractional counts distributed over the cascade members, cascade.h:286-325) and `-M 1 -! n`
    against the oracle;
  * `--crp` on the tutorial tagging cascade against the reference's own golden log (commands.trace:6976-12996): the
    sampler's per-point perplexity trajectory is RNG dependent but its level is not;
  * carmel-b200 --gpus=2 against one GPU (skipped on a one-GPU box).

Tolerances (north_star): 1e-6 relative in fp64, 1e-4 in fp32."""
import math
import os
import re

import numpy as np
import pytest

from helpers import GOLDEN, compare_wfst_text, read_history, run, stage

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli(native_lib):
    from carmel_b200 import CLI_PATH
    return CLI_PATH


def _oracle_lnp(oracle_bin, d, files, extra=()):
    from helpers import read_estimate_dump
    rc, _, err = run(oracle_bin, ["--train-cascade", *extra, f"--dump-estimate={d}/est", *files], timeout=900)
    assert rc == 0, err
    return read_estimate_dump(f"{d}/est")


# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [["--no-dense"], []])
def test_em_step_equals_estimate_then_maximize(native_lib, tmp_path, mode):
    import carmel_b200 as cb
    from carmel_b200 import synth
    w = synth.write_cipher(str(tmp_path), n_lines=40, line_len=20)
    traj = {}
    for name in ("fused", "two_call", "fused_nograph"):
        job = cb.Job(["--scaled", "-q", *mode, *w["argv"]])
        ctx = job.prepare()
        if name == "fused_nograph":
            ctx.set_option(8, 1)  # CML_OPT_NO_GRAPH
        rows = []
        for _ in range(6):
            if name == "two_call":
                r = ctx.estimate()
                d = ctx.maximize(1.0)
            else:
                r, d = ctx.em_step(1.0)
            rows.append((r.sum_ln_p, r.sum_w_ln_p, r.n_zero, d))
        traj[name] = (rows, ctx.get_params())
        job.close()
    for other in ("two_call", "fused_nograph"):
        for a, b in zip(traj["fused"][0], traj[other][0]):
            assert a[2] == b[2]
            assert abs(a[0] - b[0]) <= 1e-10 * max(1.0, abs(b[0])), (other, a, b)
            assert abs(a[3] - b[3]) <= 1e-9 * max(1.0, abs(b[3])), (other, a, b)
        wa, wb = traj["fused"][1], traj[other][1]
        fin = np.isfinite(wb)
        assert np.array_equal(fin, np.isfinite(wa))
        np.testing.assert_allclose(wa[fin], wb[fin], rtol=0, atol=1e-9)


def test_snapshot_previous_holds_the_evaluated_weights(native_lib, tmp_path):
    import carmel_b200 as cb
    from carmel_b200 import synth
    w = synth.write_cipher(str(tmp_path), n_lines=20, line_len=12)
    job = cb.Job(["--scaled", "-q", *w["argv"]])
    ctx = job.prepare()
    w0 = ctx.get_params()
    ctx.em_step(1.0)
    w1 = ctx.get_params()
    ctx.snapshot_previous(0)
    ctx.restore_params(0)
    back = ctx.get_params()
    job.close()
    fin = np.isfinite(w0)
    np.testing.assert_array_equal(back[fin], w0[fin])
    assert np.max(np.abs(w1[fin] - w0[fin])) > 0


# ----------------------------------------------------------------------------------------------------------------------
def _dev_tensor(torch, ptr, n):
    class _Arr:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}
    return torch.as_tensor(_Arr(), device="cuda")


@pytest.mark.parametrize("workload,mode", [("cipher", ["--no-dense"]), ("cipher", []), ("hmm", ["--no-dense", "--lane-min=1"]),
                                           ("hmm", [])])
def test_two_shards_through_the_kernels_equal_one(native_lib, tmp_path, workload, mode):
    """N = 2 on one GPU: rank r's context keeps block r (the product's partitioner), both E-steps run through the real
    kernels, the reduce buffers are summed in place exactly as the all-reduce would, the M-step runs on both.  Sum ln P,
    every count slot and the new parameters must equal the unsharded run."""
    import torch
    import carmel_b200 as cb
    from carmel_b200 import synth
    if workload == "cipher":
        w = synth.write_cipher(str(tmp_path), n_lines=61, line_len=17)
    else:
        w = synth.write_hmm(str(tmp_path), n_sent=333, n_tags=7, vocab=50, tags_per_word=3, len_range=(1, 30))
    one = cb.Job(["--scaled", "-q", *mode, *w["argv"]])
    c1 = one.prepare()
    noop = lambda ptr, n: None  # corpus totals are not needed for an E-step
    shards = [cb.Job(["--scaled", "-q", *mode, f"--shard={r}/2", *w["argv"]], allreduce=noop) for r in range(2)]
    ctxs = [j.prepare() for j in shards]
    assert sum(j.stats()["examples"] for j in shards) == one.stats()["examples"]
    assert sum(j.stats()["trellis_arcs"] for j in shards) == one.stats()["trellis_arcs"]
    for it in range(3):
        r1 = c1.estimate()
        counts1 = c1.counts()
        for c in ctxs:
            c.estimate_launch()
        bufs = []
        for c in ctxs:
            c.synchronize()
            p, n = c.reduce_buffer()
            bufs.append(_dev_tensor(torch, p, n))
        total = bufs[0] + bufs[1]
        for b in bufs:
            b.copy_(total)
        torch.cuda.synchronize()
        rs = [c.estimate_finish() for c in ctxs]
        for r in rs:
            assert r.n_zero == r1.n_zero
            assert abs(r.sum_ln_p - r1.sum_ln_p) <= 1e-9 * max(1.0, abs(r1.sum_ln_p)), (it, r.sum_ln_p, r1.sum_ln_p)
        for c in ctxs:
            got = c.counts()
            np.testing.assert_allclose(got, counts1, rtol=1e-9, atol=1e-300)
        c1.maximize(1.0)
        for c in ctxs:
            c.maximize(1.0)
        w1 = c1.get_params()
        fin = np.isfinite(w1)
        for c in ctxs:
            wc = c.get_params()
            assert np.array_equal(np.isfinite(wc), fin)
            np.testing.assert_allclose(wc[fin], w1[fin], rtol=0, atol=1e-9)
    for j in shards + [one]:
        j.close()


def test_empty_shard_contributes_zeros(native_lib, tmp_path):
    """a rank whose block holds no example (corpus smaller than the world) must still run its E-step"""
    import carmel_b200 as cb
    from carmel_b200 import synth
    import torch
    w = synth.write_cipher(str(tmp_path), n_lines=1, line_len=9)

    def totals_hook(ptr, n):  # stands in for the all-reduce of the corpus totals: the world has one example
        t = _dev_tensor(torch, ptr, n)
        t[0:2] = torch.clamp(t[0:2], min=1.0)
        torch.cuda.synchronize()

    jobs = [cb.Job(["--scaled", "-q", "--no-dense", f"--shard={r}/2", *w["argv"]], allreduce=totals_hook) for r in range(2)]
    ctxs = [j.prepare() for j in jobs]
    ex = [j.stats()["examples"] for j in jobs]
    assert sorted(ex) == [0, 1]
    for c, n in zip(ctxs, ex):
        r = c.estimate()
        if n == 0:
            assert r.sum_ln_p == 0.0 and r.n_zero == 0
            assert not np.any(c.counts())
        else:
            assert r.sum_ln_p < 0
    for j in jobs:
        j.close()


# ----------------------------------------------------------------------------------------------------------------------
def test_lane_kernel_selected_by_its_threshold(native_lib, oracle_bin, tmp_path):
    """16,500 narrow lattices: the lane-per-lattice kernel is chosen without --lane-min"""
    import carmel_b200 as cb
    from carmel_b200 import synth
    w = synth.write_hmm(str(tmp_path), n_sent=16500, n_tags=8, vocab=300, tags_per_word=3, len_range=(2, 12))
    est = _oracle_lnp(oracle_bin, str(tmp_path), w["files"])
    job = cb.Job(["--scaled", "--no-dense", "-q", *w["argv"]])
    ctx = job.prepare()
    assert ctx.lane_stats()["lane_examples"] == 16500
    r = ctx.estimate()
    got = ctx.example_logprob(16500)
    job.close()
    want = np.asarray(est["ln_p"])
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9)
    assert abs(r.sum_ln_p - want.sum()) <= 1e-6 * abs(want.sum())


def test_tensor_core_sweeps_selected_by_their_threshold(native_lib, oracle_bin, tmp_path):
    """16,400 fp32 sequences: the 3xTF32 dense sweeps (k_dense_tc) are chosen without CML_DENSE_TC"""
    import carmel_b200 as cb
    from carmel_b200 import synth
    assert "CML_DENSE_TC" not in os.environ
    w = synth.write_cipher(str(tmp_path), n_lines=16400, line_len=5)
    est = _oracle_lnp(oracle_bin, str(tmp_path), w["files"])
    job = cb.Job(["--float", "--scaled", "-q", *w["argv"]])
    ctx = job.prepare()
    assert ctx.dense_stats()["kernel"] == "dense_tc", ctx.dense_stats()
    ctx.estimate()
    got = ctx.example_logprob(16400)
    job.close()
    want = np.asarray(est["ln_p"])
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-6)


def test_forest_tiles_selected_by_their_threshold(native_lib, forest_oracle_bin, tmp_path):
    """8,300 forests, every one its own shape: the level-synchronous tiles (k_forest_level) are chosen without --layout"""
    import bench_forest
    from carmel_b200 import synth
    from carmel_b200.forest_api import Forests
    fs = synth.make_forests(n_forests=8300, n_rules=5000, templates=0, target_hyperedges=60)
    F = Forests(device=0, precision=64)
    F.set_rules(fs["rulespace"], fs["group_off"], fs["group_members"])
    w0 = np.full(fs["rulespace"], -np.inf)
    go, gm = fs["group_off"].astype(np.int64), fs["group_members"].astype(np.int64)
    w0[gm] = -np.log(np.repeat(np.diff(go), np.diff(go)).astype(np.float64))
    F.set_params(w0)
    F.add(fs["node_off"], fs["next"], fs["label"], fs["backref"])
    assert F.layout_stats()["tile_forests"] == 0
    lv = F.level_stats()
    assert lv["level_forests"] == 8300 and lv["small_tiles"] == lv["level_tiles"] > 100, lv
    F.estimate()
    got = F.inside(8300)[:400]
    F.close()
    p = bench_forest.inside_parity(fs, 400, 64, 0)  # (the oracle side of the same check, on the first 400 forests)
    assert p["ok"], p
    assert np.all(np.isfinite(got))


# ----------------------------------------------------------------------------------------------------------------------
def test_zero_iterations_distributes_counts_over_the_cascade(cli, oracle_bin, tmp_path):
    """-M 0 --train-cascade: <member>.trained holds the fractional counts of the member's arcs"""
    from carmel_b200 import synth
    d = str(tmp_path)
    files = {}
    for sub in ("o", "p"):
        files[sub] = synth.write_hmm(os.path.join(d, sub), n_sent=60, n_tags=5, vocab=30, tags_per_word=2, seed=11,
                                     len_range=(1, 12))["files"]
    rc, _, oerr = run(oracle_bin, ["--train-cascade", "-HJ", "-M", "0", *files["o"]])
    assert rc == 0, oerr
    rc, _, err = run(cli, ["--train-cascade", "-HJ", "-M", "0", "--scaled", "--no-dense", *files["p"]])
    assert rc == 0, err
    for n in ("tags.fsa.trained", "lexicon.fst.trained"):
        compare_wfst_text(open(os.path.join(d, "p", n)).read(), open(os.path.join(d, "o", n)).read(), 1e-6, ln_floor=-690.0)


def test_single_iteration_with_restarts_enters_the_loop(cli, tmp_path):
    """-M 1 -! 2 keeps the best of three starts (train.cc:520: the single-iteration shortcut needs ran_restarts == 0)"""
    c, f = stage(tmp_path, "epron-jpron.data", "epron-jpron.fst")
    rc, _, err = run(cli, ["-t", "-M", "1", "-!", "2", "--scaled", c, f])
    assert rc == 0, err
    assert err.count("Random restart") == 2, err


# ----------------------------------------------------------------------------------------------------------------------
# `carmel --crp -M 6000 tagging.data tagging.fsa tagging.fst` in the reference's golden log
# (carmel-tutorial/commands.trace:6976-12996): "sample prob" per-point perplexity exponents by sweep
TRACE_CRP_TAGGING = {0: 8.58505, 1: 9.03819, 2: 9.01795, 10: 8.9784, 30: 8.93824, 100: 8.90217, 300: 8.89113}


def test_crp_tagging_level_matches_the_reference_log(cli, tmp_path):
    """The sampler against the reference's own golden log.  The seed was not recorded, so the pin is distributional:
    with 24,115 points the per-point perplexity of a sweep is a tight statistic (the CPU oracle with another seed lands
    within 0.06 bits at sweep 0 and within 0.01 from sweep 10 on; tests/test_oracle_golden.py)."""
    c, a, b = stage(tmp_path, "tagging.data", "tagging.fsa", "tagging.fst")
    rc, _, err = run(cli, ["--crp", "-M", "30", "--seed=7", "--sample-prob", c, a, b], cwd=str(tmp_path), timeout=900)
    assert rc == 0, err[-2000:]
    ppx = {int(m.group(1)): float(m.group(2))
           for m in re.finditer(r"Gibbs i=(\d+) sample prob=\S+ per-point-ppx\(N=24115\)=2\^([0-9.]+)", err)}
    assert 0 in ppx and 30 in ppx, err[-1500:]
    for i, tol in ((0, 0.10), (1, 0.07), (2, 0.07), (10, 0.07), (30, 0.05)):
        assert abs(ppx[i] - TRACE_CRP_TAGGING[i]) <= tol, (i, ppx[i], TRACE_CRP_TAGGING[i])


# ----------------------------------------------------------------------------------------------------------------------
def test_cli_two_gpus_equals_one(cli, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from carmel_b200 import synth
    d = str(tmp_path)
    hist = {}
    for sub, extra in (("one", []), ("two", ["--gpus=2"])):
        w = synth.write_hmm(os.path.join(d, sub), n_sent=500, n_tags=6, vocab=40, tags_per_word=3, len_range=(1, 20))
        rc, _, err = run(cli, ["--train-cascade", "-HJ", "-M", "6", "--scaled", "--no-dense", *extra, f"--history={d}/h.{sub}",
                               *w["files"]], timeout=600)
        assert rc == 0, err
        hist[sub] = read_history(f"{d}/h.{sub}")
    assert len(hist["one"]) == len(hist["two"])
    for x, y in zip(hist["one"], hist["two"]):
        assert abs(x[1] - y[1]) <= 1e-9 * max(1.0, abs(x[1])), (x, y)
    for n in ("tags.fsa.trained", "lexicon.fst.trained"):
        compare_wfst_text(open(os.path.join(d, "two", n)).read(), open(os.path.join(d, "one", n)).read(), 1e-8, ln_floor=-690.0)


# ----------------------------------------------------------------------------------------------------------------------
def _norm_sets(path):
    txt = open(path).read()
    return sorted(tuple(sorted(int(t) for t in g.split())) for g in re.findall(r"\(([0-9 ]+)\)", txt))


def test_fem_export_matches_oracle_and_feeds_the_forest_kernels(cli, oracle_bin, tmp_path, native_lib):
    """carmel -> forest-em bridge in the PRODUCT (cascade.h:85-202, carmel.cc:756-830, to-fem.sh usage `-M -1`):
    forests byte-identical to the oracle's export, same parameters and normalisation groups; the exported cascade
    trained by forest-em-b200 follows the reference's golden cipher trajectory; --load-fem-param round trip."""
    import math
    from carmel_b200 import FOREST_CLI_PATH
    from helpers import golden
    d = str(tmp_path)
    sub = {}
    for who in ("o", "p"):
        os.makedirs(f"{d}/{who}")
        sub[who] = stage(os.path.join(d, who), "cipher.data", "cipher.wfsa", "cipher.fst")
    args = ["--train-cascade", "--normby=NC", "-HJ", "-M", "-1"]
    rc, _, err = run(oracle_bin, [*args, f"--fem-forest={d}/o.forest", f"--fem-norm={d}/o.norm", f"--fem-param={d}/o.param", *sub["o"]])
    assert rc == 0, err
    rc, _, err = run(cli, [*args, "--scaled", f"--fem-forest={d}/p.forest", f"--fem-norm={d}/p.norm", f"--fem-param={d}/p.param",
                           *sub["p"]])
    assert rc == 0, err
    assert open(f"{d}/p.forest", "rb").read() == open(f"{d}/o.forest", "rb").read()
    assert _norm_sets(f"{d}/p.norm") == _norm_sets(f"{d}/o.norm")
    from forest_helpers import read_weights
    po, pp = read_weights(f"{d}/o.param"), read_weights(f"{d}/p.param")
    assert len(po) == len(pp) == 550 + 599  # cipher.wfsa + cipher.fst arcs
    for x, y in zip(pp, po):
        assert (x == y) or abs(x - y) <= 1e-9 * max(1.0, abs(y)), (x, y)
    # the product's own export drives the forest kernels along the golden trajectory (commands.trace:6905-6950)
    want = golden()["cipher"]["trajectory_log2"]
    rc, _, err = run(FOREST_CLI_PATH, ["-U", "-f", f"{d}/p.forest", "-n", f"{d}/p.norm", "-I", f"{d}/p.param", "-i", "22", "-e", "0",
                                       f"--history={d}/h"])
    assert rc == 0, err
    hist = [(int(r[0]), float(r[1])) for r in (ln.split() for ln in open(f"{d}/h")) if r]
    assert len(hist) == len(want) == 22
    for (it, log2p), h in zip(want, hist):
        got = h[1] * 10 / math.log(2)
        assert h[0] == it and abs(got - log2p) <= 1.01e-5 * abs(log2p), (it, got, log2p)
    # --load-fem-param: the exported weights read back give the same likelihood as the transducers they came from
    rc, _, e1 = run(cli, ["--train-cascade", "--normby=NC", "-M", "-1", "--scaled", f"--history={d}/h1", *sub["p"]])
    assert rc == 0, e1
    rc, _, e2 = run(cli, ["--train-cascade", "--normby=NC", "-M", "-1", "--scaled", f"--load-fem-param={d}/p.param", f"--history={d}/h2",
                          *sub["p"]])
    assert rc == 0, e2
    a, b = read_history(f"{d}/h1"), read_history(f"{d}/h2")
    assert len(a) == len(b) == 1 and abs(a[0][1] - b[0][1]) <= 1e-9 * abs(a[0][1])


CYCLIC_FST = """qf
(q0 (q0 "a" "x" 0.5))
(q0 (q0 "a" "y" 0.2))
(q0 (q1 *e* *e* 0.3))
(q1 (q0 *e* *e* 0.4))
(q1 (q1 "a" "y" 0.3))
(q1 (qf "b" "z" 0.6))
"""


@pytest.mark.parametrize("mode,rel", [([], 1e-6), (["--scaled"], 1e-6), (["--float", "--scaled"], 1e-4)])
def test_cyclic_lattices_follow_the_reference_order(native_lib, oracle_bin, tmp_path, mode, rel):
    """An *e*:*e* loop in the transducer puts a cycle into every derivation lattice.  The reference warns ("Forward/
    backward will miss some paths", derivations.h:722-729) and walks the states once in its DFS order, back-edge
    contributions arriving late (graph.h:241-288,391-402); the oracle restates that walk.  The GPU path (k_fb_cyclic)
    must give the same likelihood trajectory and weights, beside ordinary lattices in the same batch."""
    from carmel_b200 import CLI_PATH
    from helpers import compare_wfst_text, read_history, run
    f, c = os.path.join(str(tmp_path), "c.fst"), os.path.join(str(tmp_path), "c.data")
    open(f, "w").write(CYCLIC_FST)
    open(c, "w").write('"a" "a" "b"\n"x" "y" "z"\n"a" "b"\n"y" "z"\n"b"\n"z"\n"a" "a" "a" "b"\n"y" "x" "y" "z"\n')
    rc, oout, oerr = run(oracle_bin, ["-t", "-M", "6", f"--history={tmp_path}/h.o", c, f])
    assert rc == 0, oerr
    rc, out, err = run(CLI_PATH, ["-t", "-M", "6", *mode, f"--history={tmp_path}/h.p", c, f])
    assert rc == 0, err
    assert "Warning: at least one cycle in derivations for 4 example(s)" in err
    assert "Forward/backward will miss some paths." in err
    ho, hp = read_history(f"{tmp_path}/h.o"), read_history(f"{tmp_path}/h.p")
    assert len(ho) == len(hp) and len(ho) >= 3
    for a, b in zip(hp, ho):
        assert a[0] == b[0]
        assert abs(a[1] - b[1]) <= rel * max(1.0, abs(b[1])), (a, b)
    compare_wfst_text(out, oout, rel * 20)


def _read_viterbi(path):
    out = []
    for ln in open(path):
        head = ln.split("|")[0].split()
        out.append((float(head[0]), [int(x) for x in head[2:2 + int(head[1])]]))
    return out


@pytest.mark.parametrize("seed", range(6))
def test_viterbi_random_transducers(native_lib, oracle_bin, tmp_path, seed):
    """--viterbi (cml_viterbi): the best derivation of every training pair -- weight and arc-table ids --
    equals the oracle's max-plus pass over the same lattices (random weights: no ties)"""
    from carmel_b200 import CLI_PATH
    from helpers import random_wfst, sample_pairs
    rng = np.random.default_rng(20270101 + seed)
    ns = int(rng.integers(2, 7))
    fst, ins, outs, arcs = random_wfst(rng, n_states=ns, eps_rate=float(rng.uniform(0, 0.4)))
    corpus = sample_pairs(rng, arcs, ns, n_pairs=12, ins=ins, outs=outs)
    f, c = os.path.join(str(tmp_path), "r.fst"), os.path.join(str(tmp_path), "r.data")
    open(f, "w").write(fst)
    open(c, "w").write(corpus)
    rc, _, err = run(oracle_bin, ["-t", f"--dump-viterbi={tmp_path}/v.o", c, f])
    assert rc == 0, err
    rc, _, err = run(CLI_PATH, ["-t", f"--viterbi={tmp_path}/v.p", c, f])
    assert rc == 0, err
    want, got = _read_viterbi(f"{tmp_path}/v.o"), _read_viterbi(f"{tmp_path}/v.p")
    assert len(want) == len(got) >= 5
    for (wo, po), (wp, pp) in zip(want, got):
        assert wo == wp or abs(wo - wp) <= 1e-9 * max(1.0, abs(wo)), (wo, wp)  # (-inf on both sides: only zero-weight derivations)
        if wo == -math.inf:
            continue
        # derivations that use the same arcs in another order (an insertion here or there) have the same weight: which
        # of them is reported depends on the order the states are visited in (layered order here, DFS order in the oracle)
        assert po == pp or sorted(po) == sorted(pp), (po, pp)


def test_viterbi_trained_cipher(native_lib, oracle_bin, tmp_path):
    """the tutorial's decode step on its own training data: best derivations under the TRAINED cipher channel (the
    reference's cipher.fst.trained) composed with the letter-bigram LM; also checks the member-arc labels of a path"""
    from carmel_b200 import CLI_PATH
    data, wfsa = stage(tmp_path, "cipher.data", "cipher.wfsa")
    fst = os.path.join(str(tmp_path), "cipher.fst")
    open(fst, "w").write(open(os.path.join(GOLDEN, "cipher.fst.trained")).read())
    rc, _, err = run(oracle_bin, ["--train-cascade", f"--dump-viterbi={tmp_path}/v.o", data, wfsa, fst])
    assert rc == 0, err
    rc, _, err = run(CLI_PATH, ["--train-cascade", f"--viterbi={tmp_path}/v.p", data, wfsa, fst])
    assert rc == 0, err
    assert "Viterbi: best derivations of 10 examples" in err
    want, got = _read_viterbi(f"{tmp_path}/v.o"), _read_viterbi(f"{tmp_path}/v.p")
    assert len(want) == len(got) == 10
    for (wo, po), (wp, pp) in zip(want, got):
        assert abs(wo - wp) <= 1e-9 * abs(wo), (wo, wp)
        assert po == pp
    first = open(f"{tmp_path}/v.p").readline().split("|")[1]
    assert first.count("(") == len(got[0][1])  # one group of member-arc labels per path arc
