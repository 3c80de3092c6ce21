"""Device-side lattice construction (cml_build_trellises / csrc/cml_build.cu): the GPU builder's dump
must be byte-identical to the CPU oracle's restatement of derivations::compute (derivations.h:479-704) -- state ids in
DFS pre-order after pruning, stored arc order, arc-table ids -- on the reference's fixtures, random transducers with
epsilons, random cascades, under tiny first-round capacities (every example overflows and is retried), and through
the job API at 9,000 sentences; the host builder is the second implementation beside it."""
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
                                       in_prefix="i", out_prefix="m")
    b, _, outs, arcs_b = random_wfst(rng, n_states=3, n_in=2, n_out=2, eps_rate=0.2, lock_rate=0.1, tie_rate=0.0,
                                     in_prefix="m", out_prefix="z")
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


def test_training_on_device_built_lattices(native_lib, oracle_bin, tmp_path):
    """--device-build through the job API: the E-step over GPU-built lattices gives the host-built run's likelihood and
    counts (9,000 sentences: many walks per persistent worker)."""
    import carmel_b200 as cb
    from carmel_b200 import synth
    w = synth.write_hmm(os.path.join(str(tmp_path), "h"), n_sent=9000, seed=7)
    res = {}
    for how in ("--host-build", "--device-build"):
        argv = ["-q", "--scaled", "--no-dense", how, *w["argv"]]
        job = cb.Job(argv)
        ctx = job.prepare()
        st = job.stats()
        r = ctx.estimate()
        res[how] = (st, r.sum_ln_p, ctx.counts().copy())
        job.close()
    h, g = res["--host-build"], res["--device-build"]
    assert h[0]["device_build_s"] == 0 and g[0]["device_build_s"] > 0
    assert g[0]["trellis_arcs"] == h[0]["trellis_arcs"] and g[0]["trellis_states"] == h[0]["trellis_states"]
    assert abs(g[1] - h[1]) <= 1e-12 * abs(h[1])
    np.testing.assert_allclose(g[2], h[2], rtol=1e-9, atol=1e-300)
