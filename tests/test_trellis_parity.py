"""Bit-exact trellis topology and arc-id mapping (north_star): the product's host-side lattice
builder (carmel-b200 --trellis-only, explicit-stack DFS on flat arrays, incl. the product's own
composition for cascades) against the CPU oracle's restatement of derivations::compute.
CPU only: no GPU work is involved in lattice construction."""
import filecmp
import os

import numpy as np
import pytest

from helpers import random_wfst, read_trellis_dump, run, sample_pairs, stage


@pytest.fixture(scope="session")
def cli(native_lib):
    from carmel_b200 import CLI_PATH
    assert os.path.exists(CLI_PATH)
    return CLI_PATH


def _dump_both(oracle_bin, cli, args, d):
    o, p = os.path.join(d, "oracle.trellis"), os.path.join(d, "product.trellis")
    rc, _, err = run(oracle_bin, [*args, f"--dump-trellis={o}", "--dump-estimate=/dev/null"])
    assert rc == 0, err
    rc, _, err = run(cli, [*args, "--trellis-only", f"--dump-trellis={p}"])
    assert rc == 0, err
    return o, p, err


@pytest.mark.parametrize("flags,files", [
    (["-t"], ["epron-jpron.data", "epron-jpron.fst"]),
    (["--train-cascade"], ["cipher.data", "cipher.wfsa", "cipher.fst"]),
    (["--train-cascade"], ["tagging.data", "tagging.fsa", "tagging.fst"]),
    (["-t"], ["cluster.data", "cluster.fsa"]),
    (["-t"], ["span.spell.corpus", "span.spell.wfst"]),
])
def test_reference_fixtures_bit_exact(oracle_bin, cli, tmp_path, flags, files):
    paths = stage(tmp_path, *files)
    o, p, _ = _dump_both(oracle_bin, cli, [*flags, *paths], str(tmp_path))
    assert os.path.getsize(o) > 8
    assert filecmp.cmp(o, p, shallow=False)


@pytest.mark.parametrize("seed", range(30))
def test_random_transducers_bit_exact(oracle_bin, cli, tmp_path, seed):
    rng = np.random.default_rng(20260101 + seed)
    ns = int(rng.integers(2, 7))
    fst, ins, outs, arcs = random_wfst(rng, n_states=ns, eps_rate=float(rng.uniform(0, 0.4)))
    corpus = sample_pairs(rng, arcs, ns, n_pairs=10, weighted=bool(seed % 2), ins=ins, outs=outs)
    f, c = os.path.join(str(tmp_path), "r.fst"), os.path.join(str(tmp_path), "r.data")
    open(f, "w").write(fst)
    open(c, "w").write(corpus)
    rc, _, err = run(oracle_bin, ["-t", c, f, f"--dump-trellis={tmp_path}/o", "--dump-estimate=/dev/null"])
    assert rc == 0, err
    rc, _, err = run(cli, ["-t", c, f, "--trellis-only", f"--dump-trellis={tmp_path}/p"])
    assert rc == 0, err
    t = read_trellis_dump(f"{tmp_path}/o")
    assert t["n_ex"] >= 5 and t["arc_dst"].size > 10  # the test is not vacuous
    assert filecmp.cmp(f"{tmp_path}/o", f"{tmp_path}/p", shallow=False)


@pytest.mark.parametrize("seed", range(20))
def test_random_cascades_bit_exact(oracle_bin, cli, tmp_path, seed):
    """two-member cascades: exercises the product's composition (state numbering, arc order, chains)"""
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
    rc, _, err = run(oracle_bin, ["--train-cascade", c, fa, fb, f"--dump-trellis={tmp_path}/o", "--dump-estimate=/dev/null"])
    if rc != 0:
        assert "Empty or invalid" in err or "derivation" in err, err
        return
    rc, _, err2 = run(cli, ["--train-cascade", c, fa, fb, "--trellis-only", f"--dump-trellis={tmp_path}/p"])
    assert rc == 0, err2
    assert filecmp.cmp(f"{tmp_path}/o", f"{tmp_path}/p", shallow=False)


@pytest.mark.parametrize("opt", ["--random-start", "--crp-restarts=2", "--init-em=3",
                                 "--crp-argmax-final", "--prior-inference-stddev=0.1"])
def test_unbuilt_sampler_options_are_refused(cli, tmp_path, opt):
    """Sampler options of the reference (carmel.cc:268-302) that change what is sampled and that this
    path does not build must fail loudly, before any training, rather than be ignored."""
    paths = stage(tmp_path, "cipher.data", "cipher.wfsa", "cipher.fst")
    rc, _, err = run(cli, ["--crp=2", opt, *paths])
    assert rc != 0
    assert "not implemented on the --crp path" in err


def test_multi_gpu_and_decode_options_fail_loudly(cli, tmp_path):
    """host logic that needs no GPU to be checked: the exact --crp sampler is refused with --gpus (it is sequential over
    the corpus; only --crp-batched sweeps shard), flags that change what is computed and are not built are refused, and
    the decode / device-build entry points stop with the no-CPU-fallback error on a machine without a CUDA device rather
    than computing anything on the host"""
    import torch
    paths = stage(tmp_path, "cipher.data", "cipher.wfsa", "cipher.fst")
    rc, _, err = run(cli, ["--crp", "-M", "3", "--gpus=2", *paths])
    assert rc != 0 and "sequential over the corpus" in err
    for flag in ("-r", "-a"):
        rc, _, err = run(cli, ["--train-cascade", flag, *paths])
        assert rc != 0 and "not implemented on the training path" in err
    if not torch.cuda.is_available():
        for extra in (["--viterbi=" + str(tmp_path / "v")], ["--trellis-only", "--device-build"]):
            rc, _, err = run(cli, ["--train-cascade", *extra, *paths])
            assert rc != 0 and ("no CPU fallback" in err or "CUDA" in err), err
            assert not (tmp_path / "v").exists() or (tmp_path / "v").stat().st_size == 0
