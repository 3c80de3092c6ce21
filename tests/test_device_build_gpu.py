"""Device-side lattice construction (cml_build_trellises / csrc/cml_build.cu): the GPU builder's dump
must be byte-identical to the CPU oracle's restatement of derivations::compute (derivations.h:479-704) -- state ids in
DFS pre-order after pruning, stored arc order, arc-table ids -- on the reference's fixtures, random transducers with
epsilons, random cascades, under tiny first-round capacities (every example overflows and is retried), and at a corpus
size where the product picks the GPU builder by itself; the host builder is the second implementation beside it."""
import filecmp
import os

import numpy as np
import pytest

from helpers import random_wfst, read_trellis_dump, run, sample_pairs, stage

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="session")
def cli(native_lib):
    from carmel_b200 import CLI_PATH
    assert os.path.exists(CLI_PATH)
    return CLI_PATH


def _three_way(oracle_bin, cli, args, d, env=None):
    o, h, g = (os.path.join(d, n) for n in ("oracle.trellis", "host.trellis", "device.trellis"))
    rc, _, err = run(oracle_bin, [*args, f"--dump-trellis={o}", "--dump-estimate=/dev/null"])
    if rc != 0:
        assert "Empty or invalid" in err or "derivation" in err, err
        return None
    rc, _, err = run(cli, [*args, "--trellis-only", "--host-build", f"--dump-trellis={h}"])
    assert rc == 0, err
    rc, _, err = run(cli, [*args, "--trellis-only", "--device-build", f"--dump-trellis={g}"], env=env)
    assert rc == 0, err
    assert "Device-side lattice construction" in err
    assert filecmp.cmp(o, h, shallow=False)
    assert filecmp.cmp(o, g, shallow=False), "GPU-built lattices differ from the oracle's"
    return o


@pytest.mark.parametrize("flags,files", [
    (["-t"], ["epron-jpron.data", "epron-jpron.fst"]),
    (["--train-cascade"], ["cipher.data", "cipher.wfsa", "cipher.fst"]),
    (["--train-cascade"], ["tagging.data", "tagging.fsa", "tagging.fst"]),
    (["-t"], ["cluster.data", "cluster.fsa"]),
    (["-t"], ["span.spell.corpus", "span.spell.wfst"]),
])
def test_reference_fixtures_bit_exact(oracle_bin, cli, tmp_path, flags, files):
    paths = stage(tmp_path, *files)
    o = _three_way(oracle_bin, cli, [*flags, *paths], str(tmp_path))
    assert o and os.path.getsize(o) > 8


@pytest.mark.parametrize("seed", range(12))
def test_random_transducers_bit_exact(oracle_bin, cli, tmp_path, seed):
    rng = np.random.default_rng(20260101 + seed)
    ns = int(rng.integers(2, 7))
    fst, ins, outs, arcs = random_wfst(rng, n_states=ns, eps_rate=float(rng.uniform(0, 0.4)))
    corpus = sample_pairs(rng, arcs, ns, n_pairs=10, weighted=bool(seed % 2), ins=ins, outs=outs)
    f, c = os.path.join(str(tmp_path), "r.fst"), os.path.join(str(tmp_path), "r.data")
    open(f, "w").write(fst)
    open(c, "w").write(corpus)
    # odd seeds: a first capacity round of 2 states, so every walk overflows and is retried with 4x, 16x, ... the capacity
    env = {"CML_BUILD_FIRST_STATES": "2"} if seed % 2 else None
    o = _three_way(oracle_bin, cli, ["-t", c, f], str(tmp_path), env=env)
    t = read_trellis_dump(o)
    assert t["n_ex"] >= 5 and t["arc_dst"].size > 10


@pytest.mark.parametrize("seed", range(8))
def test_random_cascades_bit_exact(oracle_bin, cli, tmp_path, seed):
    rng = np.random.default_rng(20260201 + seed)
    a, ins, mids, arcs_a = random_wfst(rng, n_states=3, n_in=2, n_out=2, eps_rate=0.2, lock_rate=0.1, tie_rate=0.0,
                                       out_prefix="m")
    b, _, outs, arcs_b = random_wfst(rng, n_states=3, n_in=2, n_out=2, eps_rate=0.2, lock_rate=0.1, tie_rate=0.0,
                                     in_syms=mids)
    lines = []
    for _ in range(24):
        li, lo = int(rng.integers(0, 4)), int(rng.integers(0, 4))
        lines.append(" ".join(ins[int(rng.integers(0, 2))] for _ in range(li)))
        lines.append(" ".join(outs[int(rng.integers(0, 2))] for _ in range(lo)))
    fa, fb, c = (os.path.join(str(tmp_path), n) for n in ("a.fst", "b.fst", "c.data"))
    open(fa, "w").write(a)
    open(fb, "w").write(b)
    open(c, "w").write("\n".join(lines) + "\n")
    _three_way(oracle_bin, cli, ["--train-cascade", c, fa, fb], str(tmp_path))


def test_auto_selection_and_training_at_scale(native_lib, oracle_bin, tmp_path):
    """>= 8,192 examples: the product builds the lattices on the GPU by itself; per-example ln P of the lattice path
    equals the oracle's on the first examples and the whole-corpus likelihood equals the host-built run's."""
    import carmel_b200 as cb
    from carmel_b200 import synth
    w = synth.write_hmm(os.path.join(str(tmp_path), "h"), n_sent=9000, seed=7)
    res = {}
    for how in ("--host-build", None):
        argv = ["-q", "--scaled", "--no-dense", *([how] if how else []), *w["argv"]]
        job = cb.Job(argv)
        ctx = job.prepare()
        st = job.stats()
        r = ctx.estimate()
        res[how] = (st, r.sum_ln_p, ctx.counts().copy())
        job.close()
    assert res["--host-build"][0]["device_build_s"] == 0
    assert res[None][0]["device_build_s"] > 0
    assert res[None][0]["trellis_arcs"] == res["--host-build"][0]["trellis_arcs"]
    assert res[None][0]["trellis_states"] == res["--host-build"][0]["trellis_states"]
    assert abs(res[None][1] - res["--host-build"][1]) <= 1e-12 * abs(res[None][1])
    np.testing.assert_allclose(res[None][2], res["--host-build"][2], rtol=1e-9, atol=1e-300)
